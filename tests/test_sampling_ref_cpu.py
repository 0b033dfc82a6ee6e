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
