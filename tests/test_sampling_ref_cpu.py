"""The sampling-filter oracle against a brute-force statement of the same definitions (CPU)."""
import numpy as np

from oracle import sampling_ref as S


def brute(logits, T, k, p):
    x = np.asarray(logits, np.float64)
    order = np.argsort(-x, kind="stable")
    keep = np.zeros(len(x), bool)
    n = len(x)
    if 0 < k < n:
        n = int((x >= x[order[k - 1]]).sum())            # ties at the k-th value stay
    cand = order[:n]
    if 0.0 < p < 1.0:
        pr = np.exp((x[cand] - x[cand].max()) / T)
        pr /= pr.sum()
        cum = np.cumsum(pr)
        m = int(np.searchsorted(cum, p - 1e-15) + 1)     # smallest head whose mass reaches p
        cand = cand[:m]
    keep[cand] = True
    return keep


def test_keep_mask_matches_brute_force():
    rng = np.random.default_rng(0)
    for trial in range(200):
        V = int(rng.integers(2, 300))
        x = rng.normal(size=V) * rng.uniform(0.5, 4)
        T = float(rng.uniform(0.2, 2.0))
        k = int(rng.integers(0, V + 2))
        p = float(rng.choice([1.0, 0.0, rng.uniform(0.05, 0.99)]))
        assert np.array_equal(S.keep_mask(x, T, k, p), brute(x, T, k, p)), (trial, V, k, p)


def test_edge_cases():
    x = np.array([5.0, 1.0, 5.0, 5.0, 0.0, 5.0, -2.0, 1.0])
    assert S.keep_mask(x, 1.0, top_k=2).sum() == 4                  # ties at the 2nd largest all stay
    assert S.keep_mask(x, 0.0, top_k=1, top_p=0.1).all()            # greedy rows are not filtered
    assert S.keep_mask(x, 1.0, top_p=1e-9).sum() >= 1
    lg = np.log(np.array([0.4, 0.3, 0.2, 0.1]))
    assert S.keep_mask(lg, 1.0, top_p=0.75).tolist() == [True, True, True, False]
    assert S.threshold(lg, 1.0, top_p=0.75) == lg[2] and S.threshold(lg, 1.0) == -np.inf


def test_penalties_and_logprob_record_against_brute_force():
    """Pins oracle/sampling_ref.penalised and logprob_record to plain-loop statements of the OpenAI definitions."""
    rng = np.random.default_rng(3)
    for _ in range(50):
        V = int(rng.integers(3, 40))
        x = rng.normal(size=V) * 3
        gen = rng.integers(0, V, size=int(rng.integers(0, 30))).tolist()
        pres, freq = float(rng.uniform(-2, 2)), float(rng.uniform(-2, 2))
        want = x.copy()
        for t in range(V):
            c = gen.count(t)
            if c:
                want[t] -= pres + freq * c
        assert np.allclose(S.penalised(x, gen, pres, freq), want, atol=1e-12)
        w = int(rng.integers(1, min(V, 21) + 1))
        t = int(rng.integers(0, V))
        ids, lps = S.logprob_record(x, t, w)
        probs = np.exp(x) / np.exp(x).sum()
        assert ids[0] == t and np.isclose(lps[0], np.log(probs[t]))
        rest = sorted(range(V), key=lambda i: (-x[i], i))[: w - 1]
        assert ids[1:].tolist() == rest and np.allclose(lps[1:], np.log(probs[rest]))
    ids, _ = S.logprob_record(np.array([1.0, 3.0, 3.0, 0.0]), 3, 3)
    assert ids.tolist() == [3, 1, 2]   # ties: lowest id first
