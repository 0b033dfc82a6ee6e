#!/usr/bin/env python
"""/v1/embeddings through the Python front the way the reference's RAG indexer calls it (one chunk per request,
`workers` concurrent requests — api/pkg/rag/rag_pgvector.go:70-83), and at a higher concurrency: chunks/s over HTTP,
request latency, how many hb_embed calls the server-side batcher needed.  One JSON line per concurrency.
  python tools/http_embed_bench.py [--model BAAI/bge-base-en-v1.5] [--tokens 512] [--requests 4000]"""
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


def client_proc(url, n_threads, per_thread, tokens, vocab, seed0, start_evt, q, per_request=1):
    import random
    u = urlparse(url)
    lat = []

    def one(i):
        rnd = random.Random(seed0 + i)
        c = http.client.HTTPConnection(u.hostname, u.port, timeout=600)
        for _ in range(per_thread):
            body = json.dumps({"input": [[rnd.randrange(vocab) for _ in range(tokens)] for _ in range(per_request)]}).encode()
            t0 = time.monotonic()
            c.request("POST", "/v1/embeddings", body, {"Content-Type": "application/json"})
            r = c.getresponse()
            d = json.loads(r.read())
            assert r.status == 200 and len(d["data"]) == per_request and len(d["data"][0]["embedding"]) > 0
            lat.append(time.monotonic() - t0)

    start_evt.wait()
    ths = [threading.Thread(target=one, args=(i,)) for i in range(n_threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    q.put(lat)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="BAAI/bge-base-en-v1.5")
    ap.add_argument("--tokens", type=int, default=512)
    ap.add_argument("--requests", type=int, default=4000)
    ap.add_argument("--workers", type=int, nargs="+", default=[10, 128, 10])
    ap.add_argument("--per-request", type=int, nargs="+", default=[1, 1, 50], help="chunks per request, one entry per --workers entry")
    a = ap.parse_args()
    from helix_b200.runtime import B200Runtime, B200RuntimeParams
    rt = B200Runtime(B200RuntimeParams(model=a.model, args=["--task", "embed", "--max-num-seqs", "256"]))
    rt.start()
    try:
        rt.warm(a.model)
        vocab = rt.engine.desc.vocab
        ctx = mp.get_context("spawn")
        for workers, per_req in zip(a.workers, a.per_request):
            procs = min(workers, 16)
            per_proc = workers // procs
            per_thread = max(1, a.requests * (4 if per_req > 1 else 1) // (procs * per_proc * per_req))
            evt, q = ctx.Event(), ctx.Queue()
            ps = [ctx.Process(target=client_proc, args=(rt.url(), per_proc, per_thread, a.tokens, vocab, 1000 * k, evt, q, per_req)) for k in range(procs)]
            for p in ps:
                p.start()
            time.sleep(3.0)
            b0 = rt.server.batcher.batches if rt.server.batcher else 0
            s0 = rt.engine.stats()
            t0 = time.monotonic()
            evt.set()
            lat = sorted(x for _ in ps for x in q.get())
            wall = time.monotonic() - t0
            s1 = rt.engine.stats()
            for p in ps:
                p.join()
            n = len(lat)
            print(json.dumps({"what": "/v1/embeddings over HTTP", "chunks_per_request": per_req, "model": a.model, "tokens_per_chunk": a.tokens,
                              "concurrent_requests": procs * per_proc, "requests": n, "wall_s": round(wall, 3),
                              "chunks_per_s": round(n * per_req / wall, 1), "latency_ms_p50": round(lat[n // 2] * 1e3, 2),
                              "latency_ms_p99": round(lat[int(n * 0.99)] * 1e3, 2),
                              "hb_embed_calls": (rt.server.batcher.batches if rt.server.batcher else 0) - b0,
                              "engine_busy_fraction": round((s1["gpu_ms_prefill"] - s0["gpu_ms_prefill"]) / 1e3 / wall, 3)}), flush=True)
    finally:
        rt.stop()


if __name__ == "__main__":
    main()
