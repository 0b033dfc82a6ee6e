"""Host-side binding of the engine C ABI (include/helix_b200.h).

Mirrors, one to one, what the Go ``B200Runtime`` shim does over cgo: create an engine
for a slot's ``gpu_index`` / ``model_memory_requirement``, load a model, start the step loop, then
submit / poll token ids (chat) or call ``embed`` (RAG ingest).
"""
import ctypes as C
from dataclasses import dataclass, field, asdict

import numpy as np

from . import _lib
from ._lib import EngineCfg, ModelDescC, SamplingC, StatsC

CAPTURE_STEP_LOGITS = 1
CAPTURE_PROMPT_LOGITS = 2


class HBError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"hb error {code}: {msg}")
        self.code = code


@dataclass
class EngineConfig:
    device: int = 0
    memory_budget_bytes: int = 0
    max_seqs: int = 256
    max_ctx: int = 8192
    max_batched_tokens: int = 16384
    kv_page_size: int = 64
    use_cuda_graphs: int = 0
    enable_prefix_cache: int = 0
    sm_budget: int = 0            # SMs this engine's persistent kernels may occupy (0 = all): multi-model packing
    sm_partition: int = 0         # 1: enforce sm_budget with a CUDA green context
    stream_priority: int = 0      # 1: high-priority stream
    decode_with_prefill: int = 0  # 1: running sequences decode inside prefill steps (mixed batches)
    fused_decode: int = 0         # 1: decode GEMMs with tile finishers instead of separate row kernels (measured slower)
    mixed_step_tokens: int = 0    # token budget of a step that carries decode rows (0 = 2048)

    def to_c(self):
        return EngineCfg(self.device, self.memory_budget_bytes, self.max_seqs, self.max_ctx, self.max_batched_tokens,
                         self.kv_page_size, self.use_cuda_graphs, self.enable_prefix_cache, self.sm_budget,
                         self.sm_partition, self.stream_priority, self.decode_with_prefill, self.fused_decode, self.mixed_step_tokens)


@dataclass
class ModelDesc:
    arch: int = 0
    hidden: int = 0
    layers: int = 0
    heads: int = 0
    kv_heads: int = 0
    head_dim: int = 0
    ffn: int = 0
    vocab: int = 0
    max_pos: int = 0
    type_vocab: int = 0
    tie_embeddings: int = 0
    norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    rope_factor: float = 0.0
    rope_low_freq_factor: float = 1.0
    rope_high_freq_factor: float = 4.0
    rope_orig_max_pos: int = 0
    qkv_bias: int = 0     # Qwen2-family decoders: biases on the q/k/v projections

    def to_c(self):
        return ModelDescC(self.arch, self.hidden, self.layers, self.heads, self.kv_heads, self.head_dim, self.ffn,
                          self.vocab, self.max_pos, self.type_vocab, self.tie_embeddings, self.norm_eps,
                          self.rope_theta, self.rope_factor, self.rope_low_freq_factor, self.rope_high_freq_factor,
                          self.rope_orig_max_pos, self.qkv_bias)


@dataclass
class Sampling:
    temperature: float = 0.0
    seed: int = 0
    max_tokens: int = 16
    eos_token: int = -1
    capture: int = 0
    top_k: int = 0      # <= 0: off
    top_p: float = 1.0  # outside (0, 1): off
    logprobs: int = 0   # n >= 1: chosen token + the n-1 most likely alternatives per generated token (n <= 21)
    presence_penalty: float = 0.0
    frequency_penalty: float = 0.0

    def to_c(self):
        return SamplingC(self.temperature, self.seed, self.max_tokens, self.eos_token, self.capture, self.top_k, self.top_p,
                         self.logprobs, self.presence_penalty, self.frequency_penalty)


def bf16_bits(a):
    """fp32 ndarray -> uint16 bf16 bit patterns (round to nearest even) — how weights cross the C ABI."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + 0x7FFF
    return ((u + r) >> 16).astype(np.uint16)


def replica_unique_id() -> bytes:
    """ncclGetUniqueId through the library (hb_replica_unique_id): 128 bytes to hand to every replica."""
    buf = C.create_string_buffer(128)
    rc = _lib.lib().hb_replica_unique_id(buf)
    if rc != 0:
        raise HBError(rc, "hb_replica_unique_id failed (libnccl.so.2 not available?)")
    return buf.raw


def gguf_describe(path) -> ModelDesc:
    d = ModelDescC()
    rc = _lib.lib().hb_gguf_describe(str(path).encode(), C.byref(d))
    if rc != 0:
        raise HBError(rc, (_lib.lib().hb_last_error(None) or b"").decode())
    return ModelDesc(**{k: getattr(d, k) for k, _ in ModelDescC._fields_ if not k.startswith("reserved")})


def gguf_read_tensor(path, hf_name):
    """One tensor of a GGUF file as fp32 [rows, cols] in the HF layout (dequantised, q/k rows un-permuted)."""
    l = _lib.lib()
    rows, cols = C.c_size_t(), C.c_size_t()
    rc = l.hb_gguf_read_tensor(str(path).encode(), hf_name.encode(), None, 0, C.byref(rows), C.byref(cols))
    if rc not in (0, -6):
        raise HBError(rc, (l.hb_last_error(None) or b"").decode() or f"tensor {hf_name} not in {path}")
    out = np.empty((rows.value, cols.value), np.float32)
    rc = l.hb_gguf_read_tensor(str(path).encode(), hf_name.encode(), out.ctypes.data, out.size, C.byref(rows), C.byref(cols))
    if rc != 0:
        raise HBError(rc, "hb_gguf_read_tensor failed")
    return out


def memory_estimate(desc: ModelDesc, cfg: EngineConfig):
    l = _lib.lib()
    w, kv, ws = C.c_uint64(), C.c_uint64(), C.c_uint64()
    rc = l.hb_memory_estimate(C.byref(desc.to_c()), C.byref(cfg.to_c()), C.byref(w), C.byref(kv), C.byref(ws))
    if rc != 0:
        raise HBError(rc, "hb_memory_estimate: unsupported description")
    return {"weights": w.value, "kv": kv.value, "workspace": ws.value}


class Engine:
    def __init__(self, cfg: EngineConfig):
        self._l = _lib.lib()
        self._h = C.c_void_p()
        self.cfg = cfg
        ccfg = cfg.to_c()
        rc = self._l.hb_engine_create(C.byref(ccfg), C.byref(self._h))
        if rc != 0:
            raise HBError(rc, (self._l.hb_last_error(None) or b"").decode())
        self.desc = None

    def _ck(self, rc):
        if rc != 0:
            raise HBError(rc, (self._l.hb_last_error(self._h) or b"").decode())

    def close(self):
        if self._h:
            self._l.hb_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- model load
    def load_random(self, desc: ModelDesc, seed=0):
        self._ck(self._l.hb_model_load_random(self._h, C.byref(desc.to_c()), seed))
        self.desc = desc

    def load_state_dict(self, desc: ModelDesc, tensors):
        """tensors: {HF checkpoint name: fp32/bf16-representable ndarray}."""
        self._ck(self._l.hb_model_load_begin(self._h, C.byref(desc.to_c())))
        for name, a in tensors.items():
            bits = bf16_bits(a)
            self._ck(self._l.hb_model_tensor_set(self._h, name.encode(), bits.ctypes.data, bits.size))
        self._ck(self._l.hb_model_load_finish(self._h))
        self.desc = desc

    def load_gguf(self, path):
        """A llama.cpp / Ollama GGUF blob: description from its metadata, tensors dequantised to bf16 (hb_model_load_gguf)."""
        self.desc = gguf_describe(path)
        self._ck(self._l.hb_model_load_gguf(self._h, str(path).encode()))

    def load_broadcast(self, desc: ModelDesc, uid: bytes, rank: int, world: int):
        """Replica load (hb_model_load_broadcast): rank 0 sends its loaded arena, the others receive it. Returns seconds."""
        sec = C.c_double()
        buf = C.create_string_buffer(bytes(uid), 128)
        self._ck(self._l.hb_model_load_broadcast(self._h, C.byref(desc.to_c()), buf, rank, world, C.byref(sec)))
        self.desc = desc
        return sec.value

    def weights_arena(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self._l.hb_model_weights_arena(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    # ---- generation
    def start(self):
        self._ck(self._l.hb_engine_start(self._h))

    def stop(self):
        self._ck(self._l.hb_engine_stop(self._h))

    def set_mixed(self, on, tokens=0):
        """decode_with_prefill / mixed_step_tokens of a live engine (next step on)."""
        self._ck(self._l.hb_engine_set_mixed(self._h, int(bool(on)), int(tokens)))

    def step(self):
        did = C.c_int()
        self._ck(self._l.hb_step(self._h, C.byref(did)))
        return bool(did.value)

    def submit(self, tokens, sampling: Sampling):
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        rid = C.c_uint64()
        sp = sampling.to_c()
        self._ck(self._l.hb_submit(self._h, t.ctypes.data, t.size, C.byref(sp), C.byref(rid)))
        return rid.value

    def poll(self, rid, cap=4096):
        buf = np.empty(cap, dtype=np.int32)
        n, fin = C.c_int32(), C.c_int32()
        self._ck(self._l.hb_poll(self._h, rid, buf.ctypes.data, cap, C.byref(n), C.byref(fin)))
        return buf[:n.value].tolist(), fin.value

    def wait(self, rid, timeout_ms=-1):
        return self._l.hb_wait(self._h, rid, timeout_ms) == 0

    def cancel(self, rid):
        self._ck(self._l.hb_cancel(self._h, rid))

    def release(self, rid):
        self._ck(self._l.hb_release(self._h, rid))

    def captured_logits(self, rid, which):
        rows = C.c_int32()
        self._ck(self._l.hb_captured_logits(self._h, rid, which, None, 0, C.byref(rows)))
        out = np.empty((rows.value, self.desc.vocab), dtype=np.float32)
        if rows.value:
            self._ck(self._l.hb_captured_logits(self._h, rid, which, out.ctypes.data, out.size, C.byref(rows)))
        return out

    def logprobs(self, rid, first_row=0, max_rows=4096):
        """(ids [rows, width], logprobs [rows, width]): column 0 = the sampled token, then the most likely tokens."""
        rows, width = C.c_int32(), C.c_int32()
        self._ck(self._l.hb_logprobs(self._h, rid, first_row, 0, None, None, C.byref(rows), C.byref(width)))
        w = max(1, width.value)
        ids = np.empty((max_rows, w), dtype=np.int32)
        lps = np.empty((max_rows, w), dtype=np.float32)
        self._ck(self._l.hb_logprobs(self._h, rid, first_row, max_rows, ids.ctypes.data, lps.ctypes.data, C.byref(rows),
                                     C.byref(width)))
        return ids[:rows.value], lps[:rows.value]

    def generate(self, prompts, sampling: Sampling):
        """Synchronous helper: drive hb_step on this thread until every prompt finished."""
        rids = [self.submit(p, sampling) for p in prompts]
        outs = [[] for _ in rids]
        done = [False] * len(rids)
        while not all(done):
            self.step()
            for i, r in enumerate(rids):
                if not done[i]:
                    toks, fin = self.poll(r)
                    outs[i] += toks
                    done[i] = fin != 0
        return rids, outs

    # ---- embeddings
    def embed(self, seqs):
        lens = [len(s) for s in seqs]
        offsets = np.zeros(len(seqs) + 1, dtype=np.int32)
        offsets[1:] = np.cumsum(lens)
        toks = np.concatenate([np.asarray(s, dtype=np.int32) for s in seqs]) if seqs else np.zeros(0, np.int32)
        return self.embed_flat(toks, offsets)

    def embed_flat(self, tokens, offsets, out=None):
        tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        n = offsets.size - 1
        if out is None:
            out = np.empty((n, self.desc.hidden), dtype=np.float32)
        self._ck(self._l.hb_embed(self._h, tokens.ctypes.data, offsets.ctypes.data, n, out.ctypes.data))
        return out

    def stats(self):
        s = StatsC()
        self._ck(self._l.hb_get_stats(self._h, C.byref(s)))
        out = {}
        for k, _ in StatsC._fields_:
            if k.startswith("reserved"):
                continue
            v = getattr(s, k)
            out[k] = list(v) if hasattr(v, "__len__") else v
        return out

    def set_profile(self, on):
        self._ck(self._l.hb_set_profile(self._h, int(bool(on))))
