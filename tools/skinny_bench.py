#!/usr/bin/env python
"""Stand-alone timing of the decode GEMM (hbk_gemm_skinny) at the Llama-3-8B decode shapes, weights rotated through
enough buffers to stay out of L2.  python tools/skinny_bench.py [M=32]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import helix_b200 as hb
from helix_b200 import _lib

L = _lib.lib()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 32
L.hbk_init()
p = lambda t: t.data_ptr()
for N, K in [(6144, 4096), (4096, 4096), (28672, 4096), (4096, 14336), (128256, 4096)]:
    nbuf = max(2, int(600e6 // (N * K * 2)) + 1)
    W = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.02 for _ in range(nbuf)]
    X = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(M, N, device="cuda")
    for w in W:
        assert L.hbk_gemm_skinny(p(X), K, p(w), K, p(out), N, M, N, K) == 0
    torch.cuda.synchronize()
    it = 40
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(it):
        L.hbk_gemm_skinny(p(X), K, p(W[i % nbuf]), K, p(out), N, M, N, K)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / it * 1e3
    print(f"skinny M={M} N={N} K={K}: {us:.1f} us  {N * K * 2 / us / 1e3:.0f} GB/s (weights)")
