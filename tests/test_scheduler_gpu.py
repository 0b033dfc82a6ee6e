"""GPU: the continuous-batching scheduler beyond FIFO whole-prompt admission — what `--max-num-seqs` under arbitrary
arrivals means in the backend the reference spawns (api/pkg/runner/vllm_runtime.go:705-762): KV pages taken on demand,
recompute preemption when the pool runs dry, and running sequences decoding inside (short) prefill steps.  Token ids are
checked against the fp32 oracle (teacher-forced: each token is the oracle's argmax or inside the near-tie margin), page
bookkeeping exactly."""
import numpy as np
import pytest

import helix_b200 as hb
from helix_b200 import configs
from oracle import weights
from oracle.llama_ref import LlamaOracle

pytestmark = pytest.mark.gpu


def tol(ref):
    return 2e-2 * max(1.0, float(np.abs(ref).max()))


def check_greedy(oracle, prompt, toks):
    oracle.reset()
    logits = oracle.forward(prompt)[-1]
    for i, t in enumerate(toks):
        best = int(np.argmax(logits))
        assert t == best or logits[best] - logits[t] <= 2 * tol(logits), f"step {i}: token {t} vs oracle {best}"
        logits = oracle.forward([t])[-1]


def drive(e, rids, max_steps=5000):
    outs = {r: [] for r in rids}
    done = set()
    steps = 0
    while len(done) < len(rids):
        e.step()
        steps += 1
        for r in rids:
            if r not in done:
                t, fin = e.poll(r)
                outs[r] += t
                if fin:
                    assert fin == 1
                    done.add(r)
        assert steps < max_steps
    return outs


def test_pages_on_demand_and_recompute_preemption():
    """A pool of 12 pages (768 positions) serves six sequences that each grow to 300 positions: pages are taken as the
    sequences grow, the youngest running sequence is preempted when the pool is empty and later resumes by re-prefilling
    prompt + generated tokens.  Every request still delivers max_tokens oracle-consistent tokens; no page is leaked."""
    d = configs.tiny_llama(layers=2, head_dim=64, vocab=1000)
    sd = weights.llama_state_dict(d, 41, 0.05)
    cfg = hb.EngineConfig(max_seqs=6, max_ctx=320, max_batched_tokens=512)
    with hb.Engine(cfg) as probe:
        probe.load_state_dict(d, sd)
        st = probe.stats()
    page_bytes = d.layers * 2 * d.kv_heads * 64 * d.head_dim * 2
    cfg.memory_budget_bytes = st["weights_bytes"] + st["workspace_bytes"] + 12 * page_bytes + page_bytes // 2
    prompts = [weights.random_tokens(300 + i, 100, d.vocab) for i in range(6)]
    with hb.Engine(cfg) as e:
        e.load_state_dict(d, sd)
        assert e.stats()["kv_pages_total"] == 12
        rids = [e.submit(p, hb.Sampling(max_tokens=200)) for p in prompts]
        outs = drive(e, rids)
        st = e.stats()
    assert st["preemptions"] > 0                                   # 6 x 5 pages never fit into 12
    assert st["kv_pages_free"] == st["kv_pages_total"] and st["running"] == 0 and st["waiting"] == 0
    o = LlamaOracle(d, sd)
    for p, r in zip(prompts, rids):
        assert len(outs[r]) == 200
        check_greedy(o, p, outs[r][:24])                           # oracle steps are O(context): a prefix of each is enough


@pytest.mark.parametrize("graphs", [0, 1])
def test_running_sequences_decode_inside_prefill_steps(graphs):
    """decode_with_prefill: while a 700-token prompt is prefilled in 128-token mixed steps, the two running sequences keep
    producing a token per step (without the option they would stall for the whole prefill), and everything generated is
    oracle-consistent — the decode rows of a mixed step run through the paged prefill-attention path."""
    d = configs.tiny_llama(layers=2, head_dim=64, vocab=1000)
    sd = weights.llama_state_dict(d, 43, 0.05)
    early = [weights.random_tokens(400 + i, n, d.vocab) for i, n in enumerate([50, 130])]
    late = weights.random_tokens(410, 700, d.vocab)
    with hb.Engine(hb.EngineConfig(max_seqs=8, max_ctx=1024, max_batched_tokens=4096, use_cuda_graphs=graphs,
                                   decode_with_prefill=1, mixed_step_tokens=128)) as e:
        e.load_state_dict(d, sd)
        r_early = [e.submit(p, hb.Sampling(max_tokens=60)) for p in early]
        got = {r: [] for r in r_early}
        for _ in range(4):                                          # prefill + a few decode steps
            e.step()
            for r in r_early:
                got[r] += e.poll(r)[0]
        before = {r: len(got[r]) for r in r_early}
        r_late = e.submit(late, hb.Sampling(max_tokens=10))
        late_out, steps_until_first = [], 0
        while not late_out:
            e.step()
            steps_until_first += 1
            for r in r_early:
                got[r] += e.poll(r)[0]
            late_out += e.poll(r_late)[0]
        during = {r: len(got[r]) - before[r] for r in r_early}
        outs = drive(e, r_early + [r_late])
        st = e.stats()
    assert steps_until_first >= 5                                   # 700 tokens in <=126-token chunks
    assert all(n >= steps_until_first - 1 for n in during.values()), (during, steps_until_first)  # one token per mixed step
    assert st["steps_mixed"] >= 5 and st["kv_pages_free"] == st["kv_pages_total"]
    o = LlamaOracle(d, sd)
    for p, r in zip(early, r_early):
        check_greedy(o, p, got[r] + outs[r])
    check_greedy(o, late, late_out + outs[r_late])
