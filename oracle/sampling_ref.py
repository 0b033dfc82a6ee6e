"""CPU oracle (test infrastructure) for the sampler's token filters — float64 numpy.

The reference forwards `top_p` (openai.ChatCompletionRequest.TopP) and vLLM's `top_k` extension untouched to its
backend (api/pkg/runner/openai_chat_handlers.go:100-175 proxies the request body); the arithmetic is the backend's.
Restated here the way vLLM 0.11 applies them (vllm/v1/sample/ops/topk_topp_sampler.py, `apply_top_k_top_p`):
top-k first — keep every logit >= the k-th largest (ties kept) — then top-p over softmax(logits/T) of the survivors:
sort ascending, drop the lowest-probability tokens whose cumulative mass is <= 1 - p, always keep the most likely token.
Parity unpinned against the reference (it has no sampling code); pinned against a brute-force statement of the same
definition in tests/test_sampling_ref_cpu.py.
"""
import numpy as np


def keep_mask(logits, temperature, top_k=0, top_p=1.0):
    """Boolean mask of the tokens that stay eligible for sampling (all True for greedy / unfiltered rows)."""
    x = np.asarray(logits, dtype=np.float64)
    V = x.shape[0]
    keep = np.ones(V, dtype=bool)
    if not temperature > 0:
        return keep
    if 0 < top_k < V:
        kth = np.partition(x, V - top_k)[V - top_k]
        keep &= x >= kth
    if 0.0 < top_p < 1.0:
        idx = np.flatnonzero(keep)
        z = np.exp((x[idx] - x[idx].max()) / temperature)
        pr = z / z.sum()
        order = np.argsort(pr, kind="stable")            # ascending
        cum = np.cumsum(pr[order])
        drop = cum <= 1.0 - top_p
        drop[-1] = False                                  # the most likely token always survives
        keep[idx[order[drop]]] = False
    return keep


def threshold(logits, temperature, top_k=0, top_p=1.0):
    """The logit value below which tokens are dropped (-inf if nothing is): what sample_threshold_kernel returns."""
    m = keep_mask(logits, temperature, top_k, top_p)
    return -np.inf if m.all() else float(np.asarray(logits, dtype=np.float64)[m].min())


def penalised(logits, generated, presence=0.0, frequency=0.0):
    """OpenAI presence / frequency penalties (the request fields openai.ChatCompletionRequest carries; the reference
    forwards them to its backend untouched, api/pkg/runner/openai_chat_handlers.go:100-175), as vLLM applies them
    (vllm/model_executor/layers/utils.py `apply_penalties`, output tokens only):
        logits[t] -= frequency * count(t in generated) + presence * [count > 0]"""
    x = np.asarray(logits, dtype=np.float64).copy()
    toks, cnt = np.unique(np.asarray(generated, dtype=np.int64), return_counts=True)
    if len(toks):
        x[toks] -= frequency * cnt + presence
    return x


def logprob_record(logits, sampled, width):
    """What hb_logprobs returns for one generated token: (ids, log-probabilities), column 0 = the sampled token, then the
    width-1 most likely tokens in descending order (lowest id first on ties); log-softmax in float64."""
    x = np.asarray(logits, dtype=np.float64)
    lse = x.max() + np.log(np.exp(x - x.max()).sum())
    order = np.lexsort((np.arange(len(x)), -x))[: max(0, width - 1)]
    ids = np.concatenate([[sampled], order]).astype(np.int64)
    return ids, x[ids] - lse
