#!/usr/bin/env python
"""How much of the engine's speed survives the Python OpenAI front (server.py) at the reference's default concurrency?
Starts B200Runtime (random-init catalogue model) in this process and drives `streams` concurrent `stream:true`
completions over HTTP from separate client PROCESSES (so the clients' JSON parsing does not share the server's GIL).
Prints one JSON line: completion tokens / wall time over HTTP, SSE chunks per second, and how busy the engine was
(device time of its forward passes / wall time) — a busy fraction near 1 means the front is not what limits it.
  python tools/http_front_bench.py [--model meta-llama/Meta-Llama-3-8B-Instruct] [--streams 256] [--prompt 256] [--decode 128]"""
import argparse
import http.client
import json
import multiprocessing as mp
import os
import sys
import threading
import time
from urllib.parse import urlparse

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def client_proc(url, n_threads, prompt_len, decode, vocab, seed0, start_evt, q):
    import random
    u = urlparse(url)
    res = []

    def one(i):
        rnd = random.Random(seed0 + i)
        body = json.dumps({"prompt": [rnd.randrange(vocab) for _ in range(prompt_len)], "max_tokens": decode, "stream": True,
                           "temperature": 0.0}).encode()
        c = http.client.HTTPConnection(u.hostname, u.port, timeout=600)
        t0 = time.monotonic()
        c.request("POST", "/v1/completions", body, {"Content-Type": "application/json"})
        r = c.getresponse()
        chunks, toks, ttft = 0, 0, None
        gaps, t_prev = [], None
        buf = b""
        while True:
            d = r.read1(65536)
            if not d:
                break
            buf += d
            while b"\n\n" in buf:
                line, buf = buf.split(b"\n\n", 1)
                if not line.startswith(b"data: ") or line == b"data: [DONE]":
                    continue
                chunks += 1
                now = time.monotonic()
                if chunks > 2:
                    gaps.append(now - t_prev)      # time between content chunks as the client sees them
                t_prev = now
                if ttft is None and chunks == 2:   # the first chunk is the empty role delta
                    ttft = now - t0
                if b'"usage"' in line:
                    toks = json.loads(line[6:]).get("usage", {}).get("completion_tokens", 0)
        res.append((r.status, chunks, toks, ttft or 0.0, time.monotonic() - t0, gaps))

    start_evt.wait()
    ths = [threading.Thread(target=one, args=(i,)) for i in range(n_threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    q.put(res)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="meta-llama/Meta-Llama-3-8B-Instruct")
    ap.add_argument("--streams", type=int, default=256)
    ap.add_argument("--procs", type=int, default=16)
    ap.add_argument("--prompt", type=int, default=256)
    ap.add_argument("--decode", type=int, default=128)
    a = ap.parse_args()
    from helix_b200.runtime import B200Runtime, B200RuntimeParams
    rt = B200Runtime(B200RuntimeParams(model=a.model, args=["--max-num-seqs", str(a.streams), "--max-model-len", str(a.prompt + a.decode + 64)]))
    rt.start()
    try:
        rt.warm(a.model)
        vocab = rt.engine.desc.vocab
        ctx = mp.get_context("spawn")
        evt, q = ctx.Event(), ctx.Queue()
        per = a.streams // a.procs
        ps = [ctx.Process(target=client_proc, args=(rt.url(), per, a.prompt, a.decode, vocab, 1000 * k, evt, q)) for k in range(a.procs)]
        for p in ps:
            p.start()
        time.sleep(3.0)  # interpreters up
        s0 = rt.engine.stats()
        t0 = time.monotonic()
        evt.set()
        res = [r for _ in ps for r in q.get()]
        wall = time.monotonic() - t0
        s1 = rt.engine.stats()
        for p in ps:
            p.join()
        ok = [r for r in res if r[0] == 200]
        toks = sum(r[2] for r in ok)
        chunks = sum(r[1] for r in ok)
        gpu_ms = (s1["gpu_ms_prefill"] + s1["gpu_ms_decode"]) - (s0["gpu_ms_prefill"] + s0["gpu_ms_decode"])
        ttfts = sorted(r[3] for r in ok)
        gaps = sorted(g for r in ok for g in r[5])
        print(json.dumps({
            "what": "completion tokens/s through the Python OpenAI front (server.py), stream:true", "model": a.model,
            "streams": a.streams, "client_processes": a.procs, "prompt_tokens": a.prompt, "decode_tokens": a.decode,
            "ok": len(ok), "completion_tokens": toks, "wall_s": round(wall, 3), "http_tokens_per_s": round(toks / wall, 1),
            "http_total_tokens_per_s": round((toks + len(ok) * a.prompt) / wall, 1),
            "sse_chunks_per_s": round(chunks / wall, 1), "tokens_per_chunk": round(toks / max(chunks, 1), 2),
            "engine_busy_fraction": round(gpu_ms / 1e3 / wall, 3),
            "engine_device_tokens_per_s": round((toks + len(ok) * a.prompt) / (gpu_ms / 1e3), 1) if gpu_ms else None,
            "chunk_gap_ms_p50": round(gaps[len(gaps) // 2] * 1e3, 2) if gaps else None,
            "chunk_gap_ms_p99": round(gaps[int(len(gaps) * 0.99)] * 1e3, 2) if gaps else None,
            "ttft_s_p50": round(ttfts[len(ttfts) // 2], 3) if ttfts else None, "ttft_s_max": round(ttfts[-1], 3) if ttfts else None,
            "decode_steps": s1["steps_decode"] - s0["steps_decode"], "mixed_steps": s1["steps_mixed"] - s0["steps_mixed"]}))
    finally:
        rt.stop()


if __name__ == "__main__":
    main()
