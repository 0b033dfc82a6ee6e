"""ctypes loader for libhelixb200.so — the same symbols a cgo shim binds."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhelixb200.so")


class LibraryMissing(RuntimeError):
    pass


class EngineCfg(C.Structure):
    _fields_ = [("device", C.c_int32), ("memory_budget_bytes", C.c_uint64), ("max_seqs", C.c_int32),
                ("max_ctx", C.c_int32), ("max_batched_tokens", C.c_int32), ("kv_page_size", C.c_int32),
                ("use_cuda_graphs", C.c_int32), ("enable_prefix_cache", C.c_int32), ("sm_budget", C.c_int32),
                ("sm_partition", C.c_int32), ("stream_priority", C.c_int32), ("decode_with_prefill", C.c_int32),
                ("fused_decode", C.c_int32), ("mixed_step_tokens", C.c_int32)]


class ModelDescC(C.Structure):
    _fields_ = [("arch", C.c_int32), ("hidden", C.c_int32), ("layers", C.c_int32), ("heads", C.c_int32),
                ("kv_heads", C.c_int32), ("head_dim", C.c_int32), ("ffn", C.c_int32), ("vocab", C.c_int32),
                ("max_pos", C.c_int32), ("type_vocab", C.c_int32), ("tie_embeddings", C.c_int32),
                ("norm_eps", C.c_float), ("rope_theta", C.c_float), ("rope_factor", C.c_float),
                ("rope_low_freq_factor", C.c_float), ("rope_high_freq_factor", C.c_float),
                ("rope_orig_max_pos", C.c_int32), ("qkv_bias", C.c_int32), ("reserved", C.c_int32 * 7)]


class SamplingC(C.Structure):
    _fields_ = [("temperature", C.c_float), ("seed", C.c_uint64), ("max_tokens", C.c_int32), ("eos_token", C.c_int32),
                ("capture", C.c_int32), ("top_k", C.c_int32), ("top_p", C.c_float), ("logprobs", C.c_int32),
                ("presence_penalty", C.c_float), ("frequency_penalty", C.c_float), ("reserved2", C.c_int32 * 4)]


class SlotInfoC(C.Structure):
    _fields_ = [("model", C.c_char * 256), ("is_embed", C.c_int32), ("tensor_parallel_size", C.c_int32),
                ("n_unknown_args", C.c_int32), ("gpu_memory_utilization", C.c_float)]


class StatsC(C.Structure):
    _fields_ = [("weights_bytes", C.c_uint64), ("kv_bytes", C.c_uint64), ("workspace_bytes", C.c_uint64),
                ("budget_bytes", C.c_uint64), ("kv_pages_total", C.c_int32), ("kv_pages_free", C.c_int32),
                ("running", C.c_int32), ("waiting", C.c_int32), ("steps_prefill", C.c_uint64),
                ("steps_decode", C.c_uint64), ("tokens_prefill", C.c_uint64), ("tokens_decode", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("graph_launches", C.c_uint64), ("cuda_error", C.c_int32),
                ("kv_pages_cached", C.c_int32), ("preemptions", C.c_int32), ("prefix_hit_tokens", C.c_uint64),
                ("steps_mixed", C.c_uint64), ("reserved", C.c_int32 * 1), ("gpu_ms_prefill", C.c_double), ("gpu_ms_decode", C.c_double),
                ("prof_ms", C.c_double * 8), ("prof_work", C.c_double * 8), ("prof_launches", C.c_uint64 * 8)]


P = C.c_void_p
I = C.c_int
# name -> (restype, argtypes): every symbol include/helix_b200.h and include/helix_b200_kernels.h declare
SIGNATURES = {
    "hb_abi_version": (I, []),
    "hb_slot_config": (I, [C.c_char_p, C.c_uint64, C.POINTER(EngineCfg), C.POINTER(SlotInfoC)]),
    "hb_engine_create": (I, [C.POINTER(EngineCfg), C.POINTER(P)]),
    "hb_engine_destroy": (None, [P]),
    "hb_last_error": (C.c_char_p, [P]),
    "hb_model_load_begin": (I, [P, C.POINTER(ModelDescC)]),
    "hb_model_tensor_set": (I, [P, C.c_char_p, P, C.c_size_t]),
    "hb_model_load_finish": (I, [P]),
    "hb_model_load_random": (I, [P, C.POINTER(ModelDescC), C.c_uint64]),
    "hb_gguf_describe": (I, [C.c_char_p, C.POINTER(ModelDescC)]),
    "hb_model_load_gguf": (I, [P, C.c_char_p]),
    "hb_gguf_read_tensor": (I, [C.c_char_p, C.c_char_p, P, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "hb_model_weights_arena": (I, [P, C.POINTER(P), C.POINTER(C.c_size_t)]),
    "hb_memory_estimate": (I, [C.POINTER(ModelDescC), C.POINTER(EngineCfg), C.POINTER(C.c_uint64),
                               C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "hb_engine_start": (I, [P]),
    "hb_engine_stop": (I, [P]),
    "hb_engine_set_mixed": (I, [P, C.c_int32, C.c_int32]),
    "hb_step": (I, [P, C.POINTER(I)]),
    "hb_submit": (I, [P, P, C.c_int32, C.POINTER(SamplingC), C.POINTER(C.c_uint64)]),
    "hb_poll": (I, [P, C.c_uint64, P, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "hb_wait": (I, [P, C.c_uint64, C.c_int32]),
    "hb_cancel": (I, [P, C.c_uint64]),
    "hb_release": (I, [P, C.c_uint64]),
    "hb_captured_logits": (I, [P, C.c_uint64, C.c_int32, P, C.c_size_t, C.POINTER(C.c_int32)]),
    "hb_logprobs": (I, [P, C.c_uint64, C.c_int32, C.c_int32, P, P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "hb_replica_unique_id": (I, [P]),
    "hb_model_load_broadcast": (I, [P, C.POINTER(ModelDescC), P, C.c_int32, C.c_int32, C.POINTER(C.c_double)]),
    "hb_embed": (I, [P, P, P, C.c_int32, P]),
    "hb_json_f32_array": (C.c_size_t, [P, C.c_size_t, P, C.c_size_t]),
    "hb_tok_load": (I, [C.c_char_p, C.POINTER(P)]),
    "hb_tok_free": (None, [P]),
    "hb_tok_vocab_size": (C.c_int32, [P]),
    "hb_tok_token_id": (C.c_int32, [P, C.c_char_p]),
    "hb_tok_encode": (I, [P, C.c_char_p, C.c_int32, P, C.c_int32, C.POINTER(C.c_int32)]),
    "hb_tok_decode": (I, [P, P, C.c_int32, C.c_int32, P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "hb_tok_chat_llama3": (I, [P, P, P, C.c_int32, P, C.c_int32, C.POINTER(C.c_int32)]),
    "hb_get_stats": (I, [P, C.POINTER(StatsC)]),
    "hb_set_profile": (I, [P, C.c_int32]),
    # kernel-level ABI
    "hbk_init": (I, []),
    "hbk_last_error": (C.c_char_p, []),
    "hbk_gemm": (I, [P, I, P, I, P, I, P, I, P, I, I, I, I, I]),
    "hbk_gemm_naive": (I, [P, I, P, I, P, I, I, I, I]),
    "hbk_gemm_skinny": (I, [P, I, P, I, P, I, I, I, I]),
    "hbk_gemm_skinny_finish": (I, [P, I, P, I, P, I, I, I, I, I]),
    "hbk_embed_gather": (I, [P, P, P, I, I]),
    "hbk_bert_embed_ln": (I, [P, P, P, P, P, P, P, P, I, I, C.c_float]),
    "hbk_rmsnorm": (I, [P, P, P, P, I, I, C.c_float]),
    "hbk_layernorm": (I, [P, P, P, P, I, I, C.c_float]),
    "hbk_rope_kv_write": (I, [P, P, P, P, P, P, I, I, I, I, I]),
    "hbk_gemm_qkv_rope": (I, [P, I, P, I, P, P, P, P, P, P, P, I, I, I, I, I, I]),
    "hbk_sample": (I, [P, I, P, P, P, I, I]),
    "hbk_sample_filtered": (I, [P, I, P, P, P, P, P, I, I]),
    "hbk_apply_penalties": (I, [P, I, P, P, I, I]),
    "hbk_logprob_topk": (I, [P, I, I, P, P, P, P, I, I]),
    "hbk_cls_pool_l2": (I, [P, P, P, I, I]),
    "hbk_attn_prefill": (I, [P, I, P, I, P, I, P, I, P, I, I, I, I, I, I, I, C.c_float]),
    "hbk_attn_prefill_paged": (I, [P, I, P, P, P, I, P, P, I, P, I, I, I, I, I, I, I, C.c_float, I]),
    "hbk_attn_naive": (I, [P, I, P, I, P, I, P, I, P, I, I, I, I, I, I, I, C.c_float]),
    "hbk_attn_decode": (I, [P, I, P, P, P, I, P, P, I, P, I, I, I, I, I, I, C.c_float, I]),
    "hbk_attn_decode_workspace_floats": (C.c_size_t, [I, I, I, I]),
}

_lib = None


def load_library(path=None):
    """Load libhelixb200.so and bind every declared symbol. Raises LibraryMissing (never falls back)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise LibraryMissing(f"{p} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "or `make -C helix_b200/csrc` (there is no CPU fallback)")
    l = C.CDLL(p, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(l, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = l
    return l


def lib():
    return load_library()
