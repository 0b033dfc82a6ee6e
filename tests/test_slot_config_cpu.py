"""CPU: hb_slot_config — the compiled decoder of the scheduler's slot-creation JSON (types.CreateRunnerSlotRequest,
api/pkg/types/runner.go:92-109) — against the Python mirror (helix_b200/runtime.py) and the reference's own known-answer
cases for --gpu-memory-utilization (api/pkg/scheduler/runner_test.go:88-186, ported in oracle/scheduler_ref.py)."""
import ctypes as C
import json

import pytest

import helix_b200 as hb
from helix_b200 import _lib, runtime as R
from oracle import scheduler_ref as S

GB = 1024 ** 3


def slot_config(obj, per_gpu=0):
    l = hb.load_library()
    cfg, info = _lib.EngineCfg(), _lib.SlotInfoC()
    rc = l.hb_slot_config(json.dumps(obj).encode() if not isinstance(obj, (bytes, str)) else (obj if isinstance(obj, bytes) else obj.encode()),
                          per_gpu, C.byref(cfg), C.byref(info))
    return rc, cfg, info


def attrs(**kw):
    a = {"runtime": "vllm", "model": "meta-llama/Meta-Llama-3-8B-Instruct"}
    a.update(kw)
    return a


def test_scheduler_emitted_slot_matches_the_python_mirror():
    # what the scheduler sends for a vLLM model: substituted args (scheduler/runner.go:1222-1259,1344-1397) inside
    # runtime_args.args, the byte budget, the GPU it chose
    for per_gpu, need in ((80 * GB, 8 * GB), (80 * GB, 1 * GB), (24 * GB, 16 * GB), (24 * GB, 20 * GB), (80 * GB, 79 * GB + 900 * 1024 * 1024)):
        args = S.substitute_vllm_args(["--max-model-len", "8192", "--gpu-memory-utilization", "{{.DynamicMemoryUtilizationRatio}}"],
                                      per_gpu, need) + ["--max-num-seqs", "64", "--trust-remote-code"]
        req = {"id": "6f1c0e0e-0000-4000-8000-000000000001",
               "attributes": attrs(model_memory_requirement=need, context_length=4096, gpu_index=3, tensor_parallel_size=1,
                                   runtime_args={"args": args})}
        rc, cfg, info = slot_config(req, per_gpu)
        assert rc == 0
        p = R.parse_vllm_args(args)
        assert cfg.device == 3 and cfg.max_seqs == p.max_num_seqs == 64 and cfg.max_ctx == p.max_model_len == 8192
        assert cfg.memory_budget_bytes == R.memory_budget(need, per_gpu, p.gpu_memory_utilization) == need   # exact bytes win
        assert abs(info.gpu_memory_utilization - p.gpu_memory_utilization) < 1e-6
        assert info.gpu_memory_utilization == pytest.approx(float(S.ratio_arg(S.vllm_memory_utilization_ratio(per_gpu, need))))
        assert info.n_unknown_args == len(p.unknown) == 1 and not info.is_embed
        assert cfg.enable_prefix_cache == 1 and cfg.decode_with_prefill == 1 and cfg.kv_page_size == 64 and cfg.use_cuda_graphs == 1
        # without the byte budget the ratio flag decides: ratio x per-GPU memory, like the mirror
        req["attributes"].pop("model_memory_requirement")
        rc, cfg, info = slot_config(req, per_gpu)
        want = R.memory_budget(0, per_gpu, p.gpu_memory_utilization)
        assert rc == 0 and abs(int(cfg.memory_budget_bytes) - want) <= per_gpu * 1e-7 + 1   # float32 ratio vs float64


def test_runtime_args_shapes_accepted_by_slot_create():
    # slot.go:410: runtime_args.model overrides; :418-447: args as strings, as mixed JSON scalars, as a {flag: value} map
    rc, cfg, info = slot_config(attrs(runtime_args={"model": "BAAI/bge-base-en-v1.5", "args": ["--task", "embed", "--max-model-len", 512]}))
    assert rc == 0 and info.model == b"BAAI/bge-base-en-v1.5" and info.is_embed == 1 and cfg.max_ctx == 512 and cfg.enable_prefix_cache == 0
    rc, cfg, info = slot_config(attrs(context_length=2048, runtime_args={"args": {"max-num-seqs": 32, "--max-num-batched-tokens": 4096.0,
                                                                                   "gpu-memory-utilization": 0.5}}), 100 * GB)
    assert rc == 0 and cfg.max_seqs == 32 and cfg.max_batched_tokens == 4096 and cfg.max_ctx == 2048   # context_length when no --max-model-len
    assert cfg.memory_budget_bytes == 50 * GB and info.model == b"meta-llama/Meta-Llama-3-8B-Instruct"
    rc, cfg, info = slot_config(attrs(runtime_args={"args": ["--no-enable-prefix-caching"]}))
    assert rc == 0 and cfg.enable_prefix_cache == 0 and cfg.max_seqs == 256 and cfg.max_ctx == 0 and cfg.device == 0  # types/memory.go:11 default
    # string escapes / unicode / nested values the decoder must skip over
    raw = ('{"attributes": {"runtime": "vllm", "model": "org/m\\u00e9 \\"q\\"", "memory_estimation_meta": {"a": [1, {"b": null}], "t": true},'
           ' "gpu_indices": [0, 1], "runtime_args": {"args": []}}, "id": "x"}')
    rc, cfg, info = slot_config(raw)
    assert rc == 0 and info.model.decode() == 'org/mé "q"'


def test_rejected_inputs():
    for bad in ('{"runtime": "ollama", "model": "m"}', '{"runtime": "vllm"}', '{"runtime": "vllm", "model": ""}', "[]", "{", '{"runtime": "vllm", "model": "m"} x',
                '{"attributes": 7}'):
        rc, _, _ = slot_config(bad)
        assert rc == -1, bad
    l = hb.load_library()
    assert l.hb_slot_config(None, 0, None, None) == -1
