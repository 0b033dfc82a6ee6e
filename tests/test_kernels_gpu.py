"""GPU parity of every sm_100a kernel, driven through the kernel-level C ABI (include/helix_b200_kernels.h)
and checked against plain fp32 torch math on the same inputs."""
import math

import numpy as np
import pytest
import torch

import helix_b200 as hb
from oracle import sampling_ref

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


@pytest.fixture(scope="module")
def L():
    l = hb.lib()
    assert l.hbk_init() == 0, l.hbk_last_error()
    return l


def ck(l, rc):
    assert rc == 0, (rc, l.hbk_last_error())
    torch.cuda.synchronize()


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(BF)


def p(t):
    return t.data_ptr() if t is not None else None


GEMM_CASES = [
    # M, N, K, epi, block_n
    (128, 64, 64, 6, 64), (128, 256, 512, 0, 256), (200, 320, 136, 0, 64), (77, 192, 72, 6, 128),
    (1024, 2048, 4096, 0, 0), (1000, 1536, 768, 1, 0), (512, 3072, 768, 2, 0), (384, 768, 3072, 4, 0),
    (640, 1024, 2048, 3, 0), (640, 1024, 512, 5, 0), (33, 1000, 256, 6, 0), (1, 512, 256, 0, 0),
    (4096, 6144, 4096, 0, 0), (300, 128256 // 8, 256, 6, 0),
]


@pytest.mark.parametrize("M,N,K,epi,bn", GEMM_CASES)
def test_gemm_epilogues(L, M, N, K, epi, bn):
    A, W = rnd(M, K, seed=1), rnd(N, K, seed=2)
    n_out = N // 2 if epi == 5 else N
    R, bias = rnd(M, n_out, scale=2.0, seed=3), rnd(N, scale=2.0, seed=4)
    acc = A.float() @ W.float().T
    if epi in (1, 2, 4):
        acc = acc + bias.float()
    if epi == 2:
        acc = torch.nn.functional.gelu(acc)
    if epi in (3, 4):
        acc = acc + R.float()
    if epi == 5:
        t = acc.view(M, N // 256, 2, 128)
        acc = (torch.nn.functional.silu(t[:, :, 0]) * t[:, :, 1]).reshape(M, N // 2)
    out = torch.full((M, n_out), float("nan"), device="cuda", dtype=torch.float32 if epi == 6 else BF)
    ck(L, L.hbk_gemm(p(A), K, p(W), K, p(out), n_out, p(R), n_out, p(bias), M, N, K, epi, bn))
    if epi == 6:
        torch.testing.assert_close(out, acc, rtol=1e-4, atol=1e-3 * math.sqrt(K / 64))
    else:
        torch.testing.assert_close(out.float(), acc, rtol=1e-2, atol=2e-2 * math.sqrt(K / 64))


def test_gemm_resid_in_place(L):
    M, N, K = 512, 1024, 1024
    A, W, X = rnd(M, K, seed=5), rnd(N, K, scale=0.05, seed=6), rnd(M, N, seed=7)
    want = X.float() + A.float() @ W.float().T
    ck(L, L.hbk_gemm(p(A), K, p(W), K, p(X), N, p(X), N, None, M, N, K, 3, 0))
    torch.testing.assert_close(X.float(), want, rtol=1e-2, atol=3e-2)


def test_gemm_linearity_property_full_size(L):
    # size-independent property at a BASELINE-sized GEMM: (A1+A2)·W == A1·W + A2·W up to fp32-accumulate rounding
    M, N, K = 8192, 4096, 4096
    A1, A2, W = rnd(M, K, seed=8), rnd(M, K, seed=9), rnd(N, K, scale=0.02, seed=10)
    A12 = (A1.float() + A2.float()).to(BF)
    outs = []
    for A in (A1, A2, A12):
        o = torch.empty(M, N, device="cuda", dtype=torch.float32)
        ck(L, L.hbk_gemm(p(A), K, p(W), K, p(o), N, None, 0, None, M, N, K, 6, 0))
        outs.append(o)
    # A12 was rounded to bf16 once: compare against the exactly-representable sum through the same kernel
    ref = torch.empty(M, N, device="cuda", dtype=torch.float32)
    ck(L, L.hbk_gemm_naive(p(A12), K, p(W), K, p(ref), N, M, N, K))
    torch.testing.assert_close(outs[2], ref, rtol=1e-4, atol=2e-3)
    assert (outs[0] + outs[1] - outs[2]).abs().max() < 0.15  # bf16 rounding of A1+A2 only


def test_rmsnorm_and_gather(L):
    T, H = 300, 4096
    x, w = rnd(T, H, seed=11), rnd(H, scale=0.1, seed=12) + 1
    out = torch.empty_like(x)
    ck(L, L.hbk_rmsnorm(p(x), p(w), p(out), None, T, H, 1e-5))
    xf = x.float()
    want = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5) * w.float()
    torch.testing.assert_close(out.float(), want, rtol=1e-2, atol=1e-2)
    idx = torch.tensor([5, 299, 0, 17], device="cuda", dtype=torch.int32)
    out2 = torch.empty(4, H, device="cuda", dtype=BF)
    ck(L, L.hbk_rmsnorm(p(x), p(w), p(out2), p(idx), 4, H, 1e-5))
    assert torch.equal(out2, out[idx.long()])
    # odd width (BERT-ish 768) and tiny width
    for H2 in (768, 64, 8192):
        x2, w2 = rnd(9, H2, seed=13), rnd(H2, seed=14)
        o2 = torch.empty_like(x2)
        ck(L, L.hbk_rmsnorm(p(x2), p(w2), p(o2), None, 9, H2, 1e-5))
        xf = x2.float()
        torch.testing.assert_close(o2.float(), xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5) * w2.float(),
                                   rtol=1e-2, atol=1e-2)


def test_layernorm_and_bert_embed(L):
    T, H, V, P = 130, 768, 1000, 512
    x, g, b = rnd(T, H, seed=15), rnd(H, scale=0.1, seed=16) + 1, rnd(H, scale=0.1, seed=17)
    out = torch.empty_like(x)
    ck(L, L.hbk_layernorm(p(x), p(g), p(b), p(out), T, H, 1e-12))
    want = torch.nn.functional.layer_norm(x.float(), (H,), g.float(), b.float(), 1e-12)
    torch.testing.assert_close(out.float(), want, rtol=1e-2, atol=1e-2)
    word, pos, typ = rnd(V, H, seed=18), rnd(P, H, seed=19), rnd(2, H, seed=20)
    tok = torch.randint(0, V, (T,), device="cuda", dtype=torch.int32)
    ps = torch.arange(T, device="cuda", dtype=torch.int32) % 100
    ck(L, L.hbk_bert_embed_ln(p(tok), p(ps), p(word), p(pos), p(typ), p(g), p(b), p(out), T, H, 1e-12))
    e = word.float()[tok.long()] + pos.float()[ps.long()] + typ.float()[0]
    torch.testing.assert_close(out.float(), torch.nn.functional.layer_norm(e, (H,), g.float(), b.float(), 1e-12),
                               rtol=1e-2, atol=1e-2)
    emb = torch.empty(T, H, device="cuda", dtype=BF)
    ck(L, L.hbk_embed_gather(p(tok), p(word), p(emb), T, H))
    assert torch.equal(emb, word[tok.long()])  # bit-exact gather


@pytest.mark.parametrize("D,Hq,Hkv", [(128, 32, 8), (64, 32, 8), (64, 4, 2)])
def test_rope_and_kv_scatter(L, D, Hq, Hkv):
    T, page, npages = 200, 64, 16
    qkv = rnd(T, (Hq + 2 * Hkv) * D, seed=21)
    orig = qkv.clone()
    pos = torch.randint(0, 3000, (T,), device="cuda", dtype=torch.int32)
    perm = torch.randperm(npages * page, device="cuda")[:T].to(torch.int32)
    perm[3] = -1  # skipped cache write
    inv = (1.0 / (500000.0 ** (torch.arange(0, D, 2, device="cuda").float() / D))).contiguous()
    kc = torch.zeros(npages, Hkv, page, D, device="cuda", dtype=BF)
    vc = torch.zeros_like(kc)
    ck(L, L.hbk_rope_kv_write(p(qkv), p(pos), p(perm), p(inv), p(kc), p(vc), T, Hq, Hkv, D, page))
    x = orig.float().view(T, Hq + 2 * Hkv, D)
    ang = pos.float()[:, None] * inv[None, :]
    cos, sin = torch.cat([ang.cos(), ang.cos()], -1)[:, None], torch.cat([ang.sin(), ang.sin()], -1)[:, None]
    qk = x[:, :Hq + Hkv]
    rot = torch.cat([-qk[..., D // 2:], qk[..., :D // 2]], -1)
    want_qk = qk * cos + rot * sin
    got = qkv.float().view(T, Hq + 2 * Hkv, D)
    torch.testing.assert_close(got[:, :Hq + Hkv], want_qk, rtol=1e-2, atol=1e-2)
    assert torch.equal(got[:, Hq + Hkv:], x[:, Hq + Hkv:])  # v untouched
    for t in range(T):
        s = int(perm[t])
        if s < 0:
            continue
        assert torch.equal(kc[s // page, :, s % page], qkv.view(T, -1, D)[t, Hq:Hq + Hkv])  # bit-exact scatter
        assert torch.equal(vc[s // page, :, s % page], qkv.view(T, -1, D)[t, Hq + Hkv:])
    assert kc.float().abs().sum(dim=(1, 3)).view(-1).count_nonzero() == T - 1


@pytest.mark.parametrize("D,Hq,Hkv,T,K,with_bias", [(128, 32, 8, 1100, 512, False), (128, 12, 2, 300, 256, True),
                                                      (64, 32, 8, 260, 256, False), (64, 4, 2, 37, 128, True)])
def test_qkv_gemm_rope_epilogue_equals_gemm_then_row_kernel(L, D, Hq, Hkv, T, K, with_bias):
    """EPI_ROPE (RoPE + KV scatter on the accumulators' way out of the QKV projection) against the two-kernel path it
    replaces: same GEMM, same bf16 rounding before the rotation, same cos/sin — the activation and both cache planes
    must be identical bit for bit (CTA-pair and single-CTA tiles, ragged last row tile, skipped slots)."""
    N, page, npages = (Hq + 2 * Hkv) * D, 64, 20
    A, W = rnd(T, K, seed=31, scale=0.5), rnd(N, K, seed=32, scale=0.5)
    bias = rnd(N, seed=33) if with_bias else None
    pos = torch.randint(0, 5000, (T,), device="cuda", dtype=torch.int32)
    slots = torch.randperm(npages * page, device="cuda")[:T].to(torch.int32)
    slots[T // 2] = -1
    inv = (1.0 / (500000.0 ** (torch.arange(0, D, 2, device="cuda").float() / D))).contiguous()
    ref = torch.empty(T, N, device="cuda", dtype=BF)
    ck(L, L.hbk_gemm(p(A), K, p(W), K, p(ref), N, None, 0, p(bias) if with_bias else None, T, N, K, 1 if with_bias else 0, 0))
    kc_ref = torch.zeros(npages, Hkv, page, D, device="cuda", dtype=BF)
    vc_ref = torch.zeros_like(kc_ref)
    ck(L, L.hbk_rope_kv_write(p(ref), p(pos), p(slots), p(inv), p(kc_ref), p(vc_ref), T, Hq, Hkv, D, page))
    got = torch.full((T, N), 7.0, device="cuda", dtype=BF)
    kc, vc = torch.zeros_like(kc_ref), torch.zeros_like(kc_ref)
    ck(L, L.hbk_gemm_qkv_rope(p(A), K, p(W), K, p(got), p(bias) if with_bias else None, p(pos), p(slots), p(inv), p(kc), p(vc),
                              T, K, Hq, Hkv, D, page))
    bad = (got != ref).nonzero()
    assert bad.numel() == 0, (f"{bad.shape[0]} mismatches, first {bad[:4].tolist()}, cols%D {sorted(set((bad[:, 1] % D).tolist()))[:16]}, "
                              f"heads {sorted(set((bad[:, 1] // D).tolist()))[:16]}, max diff {(got.float() - ref.float()).abs().max().item()}")
    assert torch.equal(kc, kc_ref) and torch.equal(vc, vc_ref)
    assert ref.float().abs().max() > 1.0  # not a vacuous comparison


def test_sampling_argmax_and_gumbel(L):
    B, V = 7, 128256
    logits = torch.randn(B, V, device="cuda")
    logits[2, 777] = logits[2, 99999] = 50.0  # tie -> lowest index
    out = torch.empty(B, device="cuda", dtype=torch.int32)
    ck(L, L.hbk_sample(p(logits), V, None, None, p(out), B, V))
    want = logits.argmax(-1)
    want[2] = 777
    assert torch.equal(out.long(), want)  # bit-exact token ids
    # temperature sampling: deterministic per seed, distribution follows softmax(logits/T)
    V2, n = 8, 4000
    lg = torch.tensor([[0.0, 1.0, 2.0, 3.0, 0.5, 1.5, 2.5, -1.0]], device="cuda").repeat(n, 1).contiguous()
    temp = torch.full((n,), 0.7, device="cuda")
    seeds = torch.arange(n, device="cuda", dtype=torch.int64) * 7919 + 13
    o1 = torch.empty(n, device="cuda", dtype=torch.int32)
    o2 = torch.empty_like(o1)
    ck(L, L.hbk_sample(p(lg), V2, p(temp), p(seeds), p(o1), n, V2))
    ck(L, L.hbk_sample(p(lg), V2, p(temp), p(seeds), p(o2), n, V2))
    assert torch.equal(o1, o2)
    freq = torch.bincount(o1.long(), minlength=V2).float() / n
    torch.testing.assert_close(freq, torch.softmax(lg[0] / 0.7, -1), atol=0.03, rtol=0)


def test_sampling_never_draws_negligible_tokens_at_full_vocab(L):
    """128256-wide rows where 64 tokens carry all but ~1e-8 of the softmax mass: over 4096 (row, seed) draws at T=1 no
    sampled id may fall outside them.  (A uniform that can reach exactly 1.0 gives +inf Gumbel noise and a uniformly
    random id once per 2^24 (token, draw) pairs, i.e. ~0.8 % of draws at this vocabulary.)"""
    n, V = 1024, 128256
    g = torch.Generator(device="cuda").manual_seed(5)
    hot = torch.randperm(V, device="cuda", generator=g)[:64]
    row = torch.full((V,), -30.0, device="cuda")
    row[hot] = torch.rand(64, device="cuda", generator=g) * 5
    lg = row.repeat(n, 1).contiguous()
    temp = torch.ones(n, device="cuda")
    ok = torch.zeros(V, dtype=torch.bool, device="cuda")
    ok[hot] = True
    seen = set()
    for rnd in range(4):
        seeds = torch.arange(n, device="cuda", dtype=torch.int64) * 104729 + 17 + rnd * 1000003
        o = torch.empty(n, device="cuda", dtype=torch.int32)
        ck(L, L.hbk_sample(p(lg), V, p(temp), p(seeds), p(o), n, V))
        assert bool(ok[o.long()].all()), o[~ok[o.long()]].tolist()
        seen.update(o.tolist())
    assert len(seen) > 32  # it does sample, not argmax


def test_penalties_and_logprobs_kernels_vs_oracle(L):
    """apply_penalties / logprob_topk at the real vocabulary width against the float64 oracle (oracle/sampling_ref.py):
    penalised logits to 1e-6, top ids bit-exact (ties -> lowest id), log-probabilities to 1e-4."""
    import numpy as np
    from oracle import sampling_ref as S
    B, V, W = 5, 128256, 21
    g = torch.Generator(device="cuda").manual_seed(9)
    logits = (torch.randn(B, V, device="cuda", generator=g) * 2).contiguous()
    logits[1, 5] = logits[1, 70000] = logits[1].max() + 1.0       # tie at the top: id 5 first
    ref = logits.cpu().double().numpy()
    rng = np.random.default_rng(4)
    gens = [rng.integers(0, V, size=n).tolist() for n in (0, 7, 300, 1, 2000)]
    gens[3] = [int(ref[3].argmax())]                               # penalise the argmax away
    pres = [0.0, 1.5, -0.5, 2.0, 0.3]
    freq = [0.0, 0.25, 0.1, 2.0, -0.2]
    off, ent = [0], []
    for b in range(B):
        toks, cnt = np.unique(np.array(gens[b], dtype=np.int64), return_counts=True)
        for t, c in zip(toks, cnt):
            ent.append((int(t), np.float32(pres[b] + freq[b] * c)))
        off.append(len(ent))
    pen = np.zeros(len(ent), dtype=[("t", np.int32), ("v", np.float32)])
    for i, (t, v) in enumerate(ent):
        pen[i] = (t, v)
    d_off = torch.tensor(off, dtype=torch.int32, device="cuda")
    d_pen = torch.from_numpy(pen.view(np.int32).reshape(-1, 2).copy()).cuda()
    ck(L, L.hbk_apply_penalties(p(logits), V, p(d_off), p(d_pen), B, V))
    torch.cuda.synchronize()
    got = logits.cpu().double().numpy()
    for b in range(B):
        want = S.penalised(ref[b], gens[b], pres[b], freq[b])
        assert np.abs(got[b] - want).max() < 1e-5, b
    sampled = torch.tensor([3, 70000, 11, int(got[3].argmax()), 99], dtype=torch.int32, device="cuda")
    width = torch.tensor([1, 21, 0, 4, 6], dtype=torch.int32, device="cuda")
    ids = torch.full((B, W), -7, dtype=torch.int32, device="cuda")
    lps = torch.full((B, W), float("nan"), device="cuda")
    ck(L, L.hbk_logprob_topk(p(logits), V, V, p(sampled), p(width), p(ids), p(lps), B, W))
    torch.cuda.synchronize()
    for b in range(B):
        w = int(width[b])
        if w == 0:
            assert int(ids[b, 0]) == -7                            # untouched row
            continue
        wi, wl = S.logprob_record(logits[b].cpu().double().numpy(), int(sampled[b]), w)
        assert ids[b, :w].tolist() == wi.tolist(), b
        assert np.abs(lps[b, :w].cpu().double().numpy() - wl).max() < 1e-4, b
    assert ids[1, 1:3].tolist() == [5, 70000]


def test_sampling_top_k_top_p(L):
    """Top-k / nucleus filtering in front of the Gumbel-max sampler: exact set membership against a float64 restatement
    (vLLM semantics: top-k first, then top-p over the survivors), untouched rows identical to the unfiltered sampler."""
    torch.manual_seed(5)
    B, V = 64, 128256
    logits = (torch.randn(1, V, device="cuda") * 2.5).repeat(B, 1).contiguous()
    temp = torch.full((B,), 0.8, device="cuda")
    seeds = torch.arange(B, device="cuda", dtype=torch.int64) * 104729 + 7
    topk = torch.zeros(B, device="cuda", dtype=torch.int32)
    topp = torch.ones(B, device="cuda")
    topk[0:16] = 1                    # == argmax whatever the noise
    topk[16:32] = 40
    topp[32:48] = 0.9
    topk[48:56], topp[48:56] = 50, 0.5
    temp[60:] = 0.0                   # greedy rows ignore the filters
    topk[60:] = 3
    out, base = torch.empty(B, device="cuda", dtype=torch.int32), torch.empty(B, device="cuda", dtype=torch.int32)
    ck(L, L.hbk_sample_filtered(p(logits), V, p(temp), p(seeds), p(topk), p(topp), p(out), B, V))
    ck(L, L.hbk_sample(p(logits), V, p(temp), p(seeds), p(base), B, V))
    out2 = torch.empty_like(out)
    ck(L, L.hbk_sample_filtered(p(logits), V, p(temp), p(seeds), p(topk), p(topp), p(out2), B, V))
    assert torch.equal(out, out2)
    row = logits[0].double().cpu().numpy()
    order = np.argsort(-row, kind="stable")
    o = out.cpu().numpy()

    def kept(top_k, top_p):  # the oracle's survivor set (a hair more inclusive: the kernel sums masses in fp32 fixed point)
        return set(np.flatnonzero(sampling_ref.keep_mask(row, 0.8, top_k, min(top_p + 1e-4, 1.0))).tolist())

    assert (o[0:16] == order[0]).all()
    assert set(o[16:32].tolist()) <= kept(40, 1.0) == set(order[:40].tolist()) and len(set(o[16:32].tolist())) > 3
    hi = kept(0, 0.9)
    assert set(o[32:48].tolist()) <= hi and 10 < len(hi) < V // 2
    assert set(o[48:56].tolist()) <= kept(50, 0.5) and len(kept(50, 0.5)) < 50
    assert (o[56:60] == base.cpu().numpy()[56:60]).all()      # unfiltered sampled rows: same token as the plain sampler
    assert (o[60:] == order[0]).all()
    # distribution over a small vocabulary: p = [.4 .3 .2 .1], top_p = 0.75 keeps {0,1,2} (0.4+0.3 < 0.75), renormalised
    n = 6000
    lg = torch.log(torch.tensor([[0.4, 0.3, 0.2, 0.1]], device="cuda")).repeat(n, 1).contiguous()
    o3 = torch.empty(n, device="cuda", dtype=torch.int32)
    sd = torch.arange(n, device="cuda", dtype=torch.int64) * 31 + 1
    t1, p75 = torch.ones(n, device="cuda"), torch.full((n,), 0.75, device="cuda")   # keep alive across the call
    ck(L, L.hbk_sample_filtered(p(lg), 4, p(t1), p(sd), None, p(p75), p(o3), n, 4))
    freq = torch.bincount(o3.long(), minlength=4).float().cpu() / n
    torch.testing.assert_close(freq, torch.tensor([0.4, 0.3, 0.2, 0.0]) / 0.9, atol=0.03, rtol=0)
    # ties at the k-th logit are all kept
    tie = torch.tensor([[5.0, 1.0, 5.0, 5.0, 0.0, 5.0, -2.0, 1.0]], device="cuda").repeat(2000, 1).contiguous()
    o4 = torch.empty(2000, device="cuda", dtype=torch.int32)
    sd4, k2 = sd[:2000].contiguous(), torch.full((2000,), 2, device="cuda", dtype=torch.int32)
    ck(L, L.hbk_sample_filtered(p(tie), 8, p(t1), p(sd4), p(k2), None, p(o4), 2000, 8))
    assert set(o4.cpu().tolist()) == {0, 2, 3, 5}


def test_cls_pool(L):
    x = rnd(50, 768, seed=30)
    first = torch.tensor([0, 7, 49], device="cuda", dtype=torch.int32)
    out = torch.empty(3, 768, device="cuda")
    ck(L, L.hbk_cls_pool_l2(p(x), p(first), p(out), 3, 768))
    torch.testing.assert_close(out, torch.nn.functional.normalize(x.float()[first.long()], dim=-1), rtol=1e-5, atol=1e-6)


def sdpa_ref(q, k, v, cu, Hq, Hkv, D, causal):
    T = q.shape[0]
    out = torch.empty(T, Hq * D, device="cuda")
    g = Hq // Hkv
    for b in range(len(cu) - 1):
        s, e = cu[b], cu[b + 1]
        Q = q[s:e].float().view(e - s, Hq, D).transpose(0, 1)
        K = k[s:e].float().view(e - s, Hkv, D).transpose(0, 1).repeat_interleave(g, 0)
        V = v[s:e].float().view(e - s, Hkv, D).transpose(0, 1).repeat_interleave(g, 0)
        o = torch.nn.functional.scaled_dot_product_attention(Q[None], K[None], V[None], is_causal=bool(causal))[0]
        out[s:e] = o.transpose(0, 1).reshape(e - s, Hq * D)
    return out


@pytest.mark.parametrize("D,Hq,Hkv,causal,lens", [
    (128, 8, 2, 1, [128]), (128, 8, 2, 1, [1, 127, 129, 300, 64]), (64, 8, 2, 1, [200, 513]),
    (64, 12, 12, 0, [512, 7, 130, 64, 1]), (128, 4, 4, 0, [257]), (128, 32, 8, 1, [2048, 1000]),
])
def test_attn_prefill_varlen(L, D, Hq, Hkv, causal, lens):
    T = sum(lens)
    cu = [0] + list(np.cumsum(lens))
    ld = (Hq + 2 * Hkv) * D
    qkv = rnd(T, ld, seed=40)  # fused layout: q | k | v column blocks of one buffer
    q, k, v = qkv[:, :Hq * D], qkv[:, Hq * D:(Hq + Hkv) * D], qkv[:, (Hq + Hkv) * D:]
    cu_t = torch.tensor(cu, device="cuda", dtype=torch.int32)
    out = torch.full((T, Hq * D), float("nan"), device="cuda", dtype=BF)
    scale = 1.0 / math.sqrt(D)
    ck(L, L.hbk_attn_prefill(p(q), ld, p(k), ld, p(v), ld, p(out), Hq * D, p(cu_t), len(lens), T, max(lens), Hq, Hkv, D,
                             causal, scale))
    want = sdpa_ref(q, k, v, cu, Hq, Hkv, D, causal)
    assert not torch.isnan(out.float()).any()
    torch.testing.assert_close(out.float(), want, rtol=2e-2, atol=2e-2)
    # second opinion: the on-device naive checker
    chk = torch.empty(T, Hq * D, device="cuda")
    ck(L, L.hbk_attn_naive(p(q), ld, p(k), ld, p(v), ld, p(chk), Hq * D, p(cu_t), len(lens), T, max(lens), Hq, Hkv, D,
                           causal, scale))
    torch.testing.assert_close(chk, want, rtol=1e-3, atol=1e-3)


def test_attn_prefill_large_logits_rescale(L):
    # scores with a strongly growing max along kv exercise the lazy O-rescale path
    D, H, n = 128, 2, 1024
    q = rnd(n, H * D, seed=41)
    k = (rnd(n, H * D, seed=42).float() * torch.linspace(0.2, 6.0, n, device="cuda")[:, None]).to(BF)
    v = rnd(n, H * D, seed=43)
    cu_t = torch.tensor([0, n], device="cuda", dtype=torch.int32)
    out = torch.empty(n, H * D, device="cuda", dtype=BF)
    ck(L, L.hbk_attn_prefill(p(q), H * D, p(k), H * D, p(v), H * D, p(out), H * D, p(cu_t), 1, n, n, H, H, D, 0,
                             1.0 / math.sqrt(D)))
    want = sdpa_ref(q, k, v, [0, n], H, H, D, 0)
    torch.testing.assert_close(out.float(), want, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("D,Hq,Hkv,splits", [(128, 32, 8, 4), (64, 32, 8, 1), (64, 4, 2, 3), (128, 8, 8, 16), (128, 16, 2, 2)])
def test_attn_decode_paged(L, D, Hq, Hkv, splits):
    page, B = 64, 5
    ctx = [1, 64, 65, 700, 2048][:B]
    max_pages = (max(ctx) + page - 1) // page
    npages = sum((c + page - 1) // page for c in ctx) + 3
    kc, vc = rnd(npages, Hkv, page, D, seed=50), rnd(npages, Hkv, page, D, seed=51)
    perm = torch.randperm(npages).tolist()
    pt = torch.zeros(B, max_pages, dtype=torch.int32)
    it = iter(perm)
    for b, c in enumerate(ctx):
        for j in range((c + page - 1) // page):
            pt[b, j] = next(it)
    q = rnd(B, Hq * D, seed=52)
    out = torch.full((B, Hq * D), float("nan"), device="cuda", dtype=BF)
    ws = torch.empty(L.hbk_attn_decode_workspace_floats(B, Hq, D, splits), device="cuda")
    ctx_t, pt_d = torch.tensor(ctx, device="cuda", dtype=torch.int32), pt.cuda()
    ck(L, L.hbk_attn_decode(p(q), Hq * D, p(kc), p(vc), p(pt_d), max_pages, p(ctx_t), p(out), Hq * D, p(ws), B, Hq, Hkv,
                            D, page, splits, 1.0 / math.sqrt(D), npages))
    g = Hq // Hkv
    for b, c in enumerate(ctx):
        pages = pt[b, :(c + page - 1) // page].long()
        K = kc[pages].permute(1, 0, 2, 3).reshape(Hkv, -1, D)[:, :c].float().repeat_interleave(g, 0)
        V = vc[pages].permute(1, 0, 2, 3).reshape(Hkv, -1, D)[:, :c].float().repeat_interleave(g, 0)
        Q = q[b].float().view(Hq, 1, D)
        s = (Q @ K.transpose(1, 2)) / math.sqrt(D)
        want = (torch.softmax(s, -1) @ V).reshape(Hq * D)
        torch.testing.assert_close(out[b].float(), want, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("D,Hq,Hkv,causal", [(128, 8, 2, 1), (64, 4, 4, 1), (128, 4, 1, 0)])
def test_attn_prefill_paged_chunks(L, D, Hq, Hkv, causal):
    """Chunked prefill: the q rows are the last q_len positions of kv_len-long sequences living in the paged pool;
    chunk starts are deliberately NOT multiples of the 128-position tile or the 64-position page."""
    page = 64
    qlen = [1, 130, 300, 256, 513, 77]
    kvlen = [1, 130, 300 + 1000, 256 + 64, 513 + 37, 77 + 2048]   # first two: whole prompt in one chunk (q_off 0)
    B = len(qlen)
    max_pages = (max(kvlen) + page - 1) // page
    npages = sum((c + page - 1) // page for c in kvlen) + 5
    kc, vc = rnd(npages, Hkv, page, D, seed=70), rnd(npages, Hkv, page, D, seed=71)
    perm = torch.randperm(npages).tolist()
    pt = torch.zeros(B, max_pages, dtype=torch.int32)
    it = iter(perm)
    for b, c in enumerate(kvlen):
        for j in range((c + page - 1) // page):
            pt[b, j] = next(it)
    T = sum(qlen)
    cu = [0]
    for n in qlen:
        cu.append(cu[-1] + n)
    q = rnd(T, Hq * D, seed=72)
    out = torch.full((T, Hq * D), float("nan"), device="cuda", dtype=BF)
    cu_t = torch.tensor(cu, device="cuda", dtype=torch.int32)
    kv_t, pt_d = torch.tensor(kvlen, device="cuda", dtype=torch.int32), pt.cuda()
    ck(L, L.hbk_attn_prefill_paged(p(q), Hq * D, p(kc), p(vc), p(pt_d), max_pages, p(kv_t), p(out), Hq * D, p(cu_t), B, T,
                                   max(qlen), Hq, Hkv, D, causal, 1.0 / math.sqrt(D), npages))
    torch.cuda.synchronize()
    g = Hq // Hkv
    for b in range(B):
        c, n = kvlen[b], qlen[b]
        pages = pt[b, :(c + page - 1) // page].long()
        K = kc[pages].permute(1, 0, 2, 3).reshape(Hkv, -1, D)[:, :c].float().repeat_interleave(g, 0)
        V = vc[pages].permute(1, 0, 2, 3).reshape(Hkv, -1, D)[:, :c].float().repeat_interleave(g, 0)
        Q = q[cu[b]:cu[b + 1]].float().view(n, Hq, D).permute(1, 0, 2)
        s = (Q @ K.transpose(1, 2)) / math.sqrt(D)
        if causal:
            qpos = torch.arange(c - n, c, device="cuda")[:, None]
            s = s.masked_fill(torch.arange(c, device="cuda")[None, :] > qpos, float("-inf"))
        want = (torch.softmax(s, -1) @ V).permute(1, 0, 2).reshape(n, Hq * D)
        torch.testing.assert_close(out[cu[b]:cu[b + 1]].float(), want, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("M,N,K", [(32, 6144, 4096), (1, 4096, 4096), (7, 1000, 256), (64, 4096, 14336), (100, 28672, 4096),
                                   (256, 128256, 2048), (200, 512, 64), (33, 128, 4096)])
def test_gemm_skinny_stream_k(L, M, N, K):
    X, W = rnd(M, K, seed=60), rnd(N, K, scale=0.05, seed=61)
    out = torch.full((M, N), float("nan"), device="cuda")
    ck(L, L.hbk_gemm_skinny(p(X), K, p(W), K, p(out), N, M, N, K))
    ref = X.float() @ W.float().T  # fp32 torch math on the same bf16 inputs (TF32 is off by default for matmul)
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=2e-3)
    out2 = torch.empty_like(out)
    ck(L, L.hbk_gemm_skinny(p(X), K, p(W), K, p(out2), N, M, N, K))
    assert torch.equal(out, out2)  # slab order is fixed: bit-deterministic (no atomics)
    # fused tile finisher (the decode step's path): the last-arriving CTA of each tile sums the slabs in the same fixed
    # order, so the result is bit-identical to the separate sum pass — on every one of three back-to-back launches
    out3 = torch.full((M, N), float("nan"), device="cuda")
    ck(L, L.hbk_gemm_skinny_finish(p(X), K, p(W), K, p(out3), N, M, N, K, 3))
    assert torch.equal(out, out3)
