#!/usr/bin/env python
"""BASELINE.json configs[3]: Llama-3-8B + Llama-3.2-1B + bge-base co-resident on ONE B200 under the scheduler's
memory-fit rule, mixed chat + embed traffic; reports each model's throughput solo and packed (interference).

    python tools/pack_bench.py > profiles/r01_pack.json
"""
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import helix_b200 as hb  # noqa: E402
from helix_b200 import configs  # noqa: E402
from helix_b200.engine import memory_estimate  # noqa: E402

GB = 1024 ** 3


def chat_load(e, desc, sessions, prompt, decode, seconds, out, key):
    rng = np.random.default_rng(0)
    prompts = [rng.integers(0, desc.vocab, size=prompt).astype(np.int32) for _ in range(sessions)]
    toks, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        rids = [e.submit(p, hb.Sampling(max_tokens=decode)) for p in prompts]
        for r in rids:
            fin = 0
            while not fin:
                e.wait(r, 60000)
                t, fin = e.poll(r)
                toks += len(t)
            e.release(r)
        toks += sessions * prompt
    out[key] = toks / (time.perf_counter() - t0)


def embed_load(e, desc, chunks, seconds, out, key):
    rng = np.random.default_rng(1)
    toks = rng.integers(0, desc.vocab, size=chunks * 512).astype(np.int32)
    offs = (np.arange(chunks + 1) * 512).astype(np.int32)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        e.embed_flat(toks, offs)
        n += chunks
    out[key] = n / (time.perf_counter() - t0)


def main():
    seconds = float(os.environ.get("PACK_SECONDS", "8"))
    specs = [("llama3_8b", configs.llama3_8b(), hb.EngineConfig(max_seqs=16, max_ctx=2304, max_batched_tokens=8192, use_cuda_graphs=1)),
             ("llama32_1b", configs.llama32_1b(), hb.EngineConfig(max_seqs=16, max_ctx=2304, max_batched_tokens=8192, use_cuda_graphs=1)),
             ("bge_base", configs.bge_base(), hb.EngineConfig(max_seqs=64, max_ctx=512, max_batched_tokens=32768))]
    total = 183359 * 1024 * 1024
    allocated, budgets = 0, {}
    for name, d, cfg in specs:
        est = memory_estimate(d, cfg)
        need = sum(est.values()) + (256 << 20)
        # the reference's single-GPU fit rule (api/pkg/scheduler/global_allocator.go:349-452): total - allocated >= need
        assert total - allocated >= need, "scheduler would not place this slot"
        budgets[name] = need
        allocated += need
        cfg.memory_budget_bytes = need
    engines = {}
    for name, d, cfg in specs:
        e = hb.Engine(cfg)
        e.load_random(d, 1)
        if d.arch == configs.LLAMA:
            e.start()
        engines[name] = (e, d)
    res = {"budgets_gb": {k: round(v / GB, 2) for k, v in budgets.items()}, "gpu_total_gb": round(total / GB, 1)}

    def loads(out):
        return [threading.Thread(target=chat_load, args=(engines["llama3_8b"][0], engines["llama3_8b"][1], 16, 2048, 64, seconds, out, "llama3_8b_tok_s")),
                threading.Thread(target=chat_load, args=(engines["llama32_1b"][0], engines["llama32_1b"][1], 16, 2048, 64, seconds, out, "llama32_1b_tok_s")),
                threading.Thread(target=embed_load, args=(engines["bge_base"][0], engines["bge_base"][1], 4096, seconds, out, "bge_chunks_s"))]
    solo = {}
    for t in loads(solo):
        t.start()
        t.join()
    packed = {}
    ts = loads(packed)
    [t.start() for t in ts]
    [t.join() for t in ts]
    res["solo"], res["packed"] = solo, packed
    res["packed_over_solo"] = {k: packed[k] / solo[k] for k in solo}
    for name, (e, d) in engines.items():
        st = e.stats()
        assert st["weights_bytes"] + st["kv_bytes"] + st["workspace_bytes"] <= st["budget_bytes"]
        e.close()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
