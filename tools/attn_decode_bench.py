#!/usr/bin/env python
"""Standalone timing of the decode-attention kernel (L8B head shape) over batch sizes / split counts: does the kernel's
bandwidth depend on how many of the 296 CTA slots (2 per SM) the grid fills?
  python tools/attn_decode_bench.py [ctx=2048]"""
import ctypes as C
import math
import sys

import torch

sys.path.insert(0, ".")
from helix_b200 import _lib  # noqa: E402

L = _lib.lib()
assert L.hbk_init() == 0
ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
Hq, Hkv, D, page = 32, 8, 128, 64
BF = torch.bfloat16


def p(t):
    return t.data_ptr()


for B, splits in [(32, 1), (37, 1), (36, 1), (24, 1), (18, 2), (16, 2), (32, 2), (64, 1), (74, 1), (148, 1)]:
    pages_per = (ctx + page - 1) // page
    npages = B * pages_per + 1
    kc = (torch.randn(npages, Hkv, page, D, device="cuda") * 0.5).to(BF)
    vc = (torch.randn(npages, Hkv, page, D, device="cuda") * 0.5).to(BF)
    pt = torch.randperm(npages - 1, device="cuda").to(torch.int32).view(B, pages_per).contiguous()
    q = torch.randn(B, Hq * D, device="cuda").to(BF)
    out = torch.empty(B, Hq * D, device="cuda", dtype=BF)
    ws = torch.empty(L.hbk_attn_decode_workspace_floats(B, Hq, D, max(splits, 1)), device="cuda")
    ctx_t = torch.full((B,), ctx, device="cuda", dtype=torch.int32)
    flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)

    def run():
        rc = L.hbk_attn_decode(p(q), Hq * D, p(kc), p(vc), p(pt), pages_per, p(ctx_t), p(out), Hq * D, p(ws), B, Hq, Hkv, D, page,
                               splits, 1.0 / math.sqrt(D), npages)
        assert rc == 0, rc

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()  # KV of one call (<= 600 MB) mostly exceeds L2 anyway; flush for the small batches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    gb = B * ctx * 2 * Hkv * D * 2 / 1e9
    print(f"B={B:4d} splits={splits} ctas={B * Hkv * splits:5d} : {ms * 1e3:8.1f} us  {gb / (ms * 1e-3):7.0f} GB/s")
