"""GPU parity of the whole hot path through the engine C ABI (include/helix_b200.h) against the oracle
and the committed HF golden fixtures.  Tolerance (bf16 activations vs fp32 oracle, stated per north_star):
per-token logit max-abs-diff <= 2e-2 * max(1, ||logits||_inf) (the measured worst case is
about a third of it, see tests/test_baseline_shapes_gpu.py which prints and ratchets the ratio); token ids bit-exact wherever the oracle's
top-1 margin exceeds twice that bound."""
import os

import numpy as np
import pytest

import helix_b200 as hb
from helix_b200 import configs
from helix_b200.engine import CAPTURE_PROMPT_LOGITS, CAPTURE_STEP_LOGITS
from oracle import weights
from oracle.bert_ref import bert_embed
from oracle.llama_ref import LlamaOracle

pytestmark = pytest.mark.gpu

LLAMA_CASES = {
    "llama_tiny_d64": lambda: configs.tiny_llama(layers=2, head_dim=64, vocab=1000),
    "llama_tiny_d64_s05": lambda: configs.tiny_llama(layers=2, head_dim=64, vocab=1000),
    "llama_tiny_d128_rope3": lambda: configs.tiny_llama(layers=3, head_dim=128, vocab=1000, rope_scaling=True),
}


def tol(ref):
    return 2e-2 * max(1.0, float(np.abs(ref).max()))


def check_tokens_against(oracle, prompt, toks, rows):
    """Teacher-forced check: at every step the engine's token must be the oracle argmax unless the margin is a near-tie."""
    oracle.reset()
    logits = oracle.forward(prompt)[-1]
    for i, t in enumerate(toks):
        bound = tol(logits)
        assert np.abs(rows[i] - logits).max() <= bound, f"step {i}: logit diff {np.abs(rows[i] - logits).max()}"
        best = int(np.argmax(logits))
        if t != best:
            assert logits[best] - logits[t] <= 2 * bound, f"step {i}: token {t} vs oracle {best}"
        logits = oracle.forward([t])[-1]


def check_greedy(oracle, prompt, toks):
    """Tokens only: every generated token is the oracle's argmax, or within the near-tie margin of it."""
    oracle.reset()
    logits = oracle.forward(prompt)[-1]
    for i, t in enumerate(toks):
        best = int(np.argmax(logits))
        assert t == best or logits[best] - logits[t] <= 2 * tol(logits), f"step {i}: token {t} vs oracle {best}"
        logits = oracle.forward([t])[-1]


@pytest.mark.parametrize("name", sorted(LLAMA_CASES))
@pytest.mark.parametrize("graphs,fused", [(0, 0), (1, 0), (1, 1)])
def test_llama_matches_golden_and_oracle(name, graphs, fused, golden_dir):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    d = LLAMA_CASES[name]()
    sd = weights.llama_state_dict(d, int(g["seed"]), float(g["std"]))
    prompt = g["prompt"]
    n_dec = len(g["greedy_tokens"])
    # fused = 1: the decode GEMMs' tile finishers (RoPE + KV write, residual + norm statistics, SwiGLU inside the GEMM)
    with hb.Engine(hb.EngineConfig(max_seqs=8, max_ctx=512, max_batched_tokens=1024, use_cuda_graphs=graphs, fused_decode=fused)) as e:
        e.load_state_dict(d, sd)
        rids, outs = e.generate([prompt], hb.Sampling(max_tokens=n_dec, capture=CAPTURE_PROMPT_LOGITS | CAPTURE_STEP_LOGITS))
        pl = e.captured_logits(rids[0], CAPTURE_PROMPT_LOGITS)
        sl = e.captured_logits(rids[0], CAPTURE_STEP_LOGITS)
        st = e.stats()
    assert pl.shape == g["prompt_logits"].shape and sl.shape == g["step_logits"].shape
    assert np.abs(pl - g["prompt_logits"]).max() <= tol(g["prompt_logits"])  # vs HF fixture
    assert np.abs(sl[0] - pl[-1]).max() < 1e-5                               # last-position path == all-position path
    check_tokens_against(LlamaOracle(d, sd), prompt, outs[0], sl)
    if name == "llama_tiny_d64":  # comfortable margins (>=0.14): token ids must be bit-exact vs HF
        assert outs[0] == g["greedy_tokens"].tolist()
    assert st["kv_pages_free"] == st["kv_pages_total"] and st["running"] == 0  # page bookkeeping exact
    assert st["kernel_launches"] > 0 and st["cuda_error"] == 0
    if graphs:
        assert st["graph_launches"] == n_dec - 1


def test_continuous_batching_matches_solo_runs():
    """Ragged prompts admitted at different steps (max_seqs smaller than the request count) must produce exactly
    what each request produces alone: batching is invisible (bit-exact token ids)."""
    d = configs.tiny_llama(layers=2, head_dim=64, vocab=1000)
    sd = weights.llama_state_dict(d, 7, 0.05)
    lens = [1, 5, 63, 64, 65, 130, 257, 31, 400, 2]
    prompts = [weights.random_tokens(100 + i, n, d.vocab) for i, n in enumerate(lens)]
    max_new = [3, 9, 17, 2, 12, 1, 20, 5, 7, 11]
    solo = []
    with hb.Engine(hb.EngineConfig(max_seqs=4, max_ctx=512, max_batched_tokens=512)) as e:
        e.load_state_dict(d, sd)
        for pr, m in zip(prompts, max_new):
            _, o = e.generate([pr], hb.Sampling(max_tokens=m))
            solo.append(o[0])
        # all at once: admission is limited by max_seqs=4 and the 512-token prefill budget
        rids = [e.submit(pr, hb.Sampling(max_tokens=m)) for pr, m in zip(prompts, max_new)]
        outs = [[] for _ in rids]
        done = [False] * len(rids)
        steps = 0
        while not all(done):
            e.step()
            steps += 1
            assert e.stats()["running"] <= 4
            for i, r in enumerate(rids):
                if not done[i]:
                    t, fin = e.poll(r)
                    outs[i] += t
                    done[i] = fin != 0
            assert steps < 500
        st = e.stats()
    assert [len(o) for o in outs] == max_new
    assert outs == solo
    assert st["kv_pages_free"] == st["kv_pages_total"]


@pytest.mark.parametrize("head_dim,budget", [(64, 100), (128, 256)])
def test_chunked_prefill_of_prompts_longer_than_the_step_budget(head_dim, budget):
    """A prompt longer than max_batched_tokens is prefilled in budget-sized chunks (chunks after the first attend to the
    cached prefix through the paged pool; budget 100 puts chunk starts off every tile/page boundary).  Every prompt
    position's logits and the decoded tokens must match the fp32 oracle, and short prompts queued behind the long one
    must still produce exactly their solo outputs."""
    d = configs.tiny_llama(layers=2, head_dim=head_dim, vocab=1000)
    sd = weights.llama_state_dict(d, 11, 0.05)
    long_prompt = weights.random_tokens(900, 3 * budget + 57, d.vocab)
    shorts = [weights.random_tokens(901 + i, n, d.vocab) for i, n in enumerate([5, 70, budget])]
    cap = CAPTURE_PROMPT_LOGITS | CAPTURE_STEP_LOGITS
    with hb.Engine(hb.EngineConfig(max_seqs=8, max_ctx=1024, max_batched_tokens=budget)) as e:
        e.load_state_dict(d, sd)
        solo = [e.generate([pr], hb.Sampling(max_tokens=6))[1][0] for pr in shorts]
        before = e.stats()["steps_prefill"]
        rid = e.submit(long_prompt, hb.Sampling(max_tokens=8, capture=cap))
        rs = [e.submit(pr, hb.Sampling(max_tokens=6)) for pr in shorts]
        outs = {r: [] for r in [rid] + rs}
        done = set()
        steps = 0
        while len(done) < len(outs):
            e.step()
            steps += 1
            for r in outs:
                if r not in done:
                    t, fin = e.poll(r)
                    outs[r] += t
                    if fin:
                        done.add(r)
            assert steps < 200
        pl = e.captured_logits(rid, CAPTURE_PROMPT_LOGITS)
        sl = e.captured_logits(rid, CAPTURE_STEP_LOGITS)
        st = e.stats()
    assert st["steps_prefill"] - before >= 4                      # 3 full chunks + the tail (+ the short prompts)
    oracle = LlamaOracle(d, sd)
    want = oracle.forward(long_prompt)
    assert pl.shape == want.shape
    assert np.abs(pl - want).max() <= tol(want)
    check_tokens_against(oracle, long_prompt, outs[rid], sl)
    assert [outs[r] for r in rs] == solo
    assert st["kv_pages_free"] == st["kv_pages_total"]


def test_top_k_top_p_through_the_engine():
    """hb_sampling.top_k / top_p reach the sampler in prefill steps and in (graph-captured) decode steps: top_k=1 at any
    temperature is greedy decoding; a mixed batch leaves the unfiltered request's tokens unchanged; seeded runs repeat."""
    d = configs.tiny_llama(layers=2, head_dim=64, vocab=1000)
    sd = weights.llama_state_dict(d, 3, 0.05)
    pr = [weights.random_tokens(40 + i, 20 + 7 * i, d.vocab) for i in range(3)]
    with hb.Engine(hb.EngineConfig(max_seqs=4, max_ctx=256, max_batched_tokens=256, use_cuda_graphs=1)) as e:
        e.load_state_dict(d, sd)
        greedy = e.generate(pr, hb.Sampling(max_tokens=12))[1]
        k1 = e.generate(pr, hb.Sampling(max_tokens=12, temperature=1.5, seed=9, top_k=1))[1]
        assert k1 == greedy
        plain = e.generate([pr[0]], hb.Sampling(max_tokens=12, temperature=0.9, seed=4))[1][0]
        ra = e.submit(pr[0], hb.Sampling(max_tokens=12, temperature=0.9, seed=4))
        rb = e.submit(pr[1], hb.Sampling(max_tokens=12, temperature=0.9, seed=5, top_p=0.3, top_k=20))
        outs = {ra: [], rb: []}
        done = set()
        while len(done) < 2:
            e.step()
            for r in outs:
                t, fin = e.poll(r)
                outs[r] += t
                if fin:
                    done.add(r)
        assert outs[ra] == plain
        again = e.generate([pr[1]], hb.Sampling(max_tokens=12, temperature=0.9, seed=5, top_p=0.3, top_k=20))[1][0]
        assert again == outs[rb] and len(again) == 12
        tiny_p = e.generate(pr, hb.Sampling(max_tokens=12, temperature=2.0, seed=1, top_p=1e-6))[1]
        assert tiny_p == greedy                      # a vanishing nucleus keeps only the most likely token


def test_prefix_cache_reuses_pages_across_requests_and_turns():
    """enable_prefix_cache: leading full pages (64 tokens) of a prompt that are already in the pool are not prefilled
    again — a shared system prompt, and the next turn of a chat (which resends prompt + generated tokens).  The rest of
    the prompt runs through the paged chunk path; logits/tokens still match the fp32 oracle; nothing leaks; an
    over-committed pool evicts unreferenced cached pages instead of refusing work."""
    d = configs.tiny_llama(layers=2, head_dim=64, vocab=1000)
    sd = weights.llama_state_dict(d, 21, 0.05)
    oracle = LlamaOracle(d, sd)
    sys_p = weights.random_tokens(300, 200, d.vocab)
    p1 = np.concatenate([sys_p, weights.random_tokens(301, 50, d.vocab)])
    p2 = np.concatenate([sys_p, weights.random_tokens(302, 70, d.vocab)])
    cap = CAPTURE_STEP_LOGITS
    with hb.Engine(hb.EngineConfig(max_seqs=4, max_ctx=512, max_batched_tokens=512, enable_prefix_cache=1)) as e:
        e.load_state_dict(d, sd)
        r1, o1 = e.generate([p1], hb.Sampling(max_tokens=20, capture=cap))
        s1 = e.stats()
        assert s1["prefix_hit_tokens"] == 0 and s1["kv_pages_cached"] == (250 + 19) // 64
        r2, o2 = e.generate([p2], hb.Sampling(max_tokens=8, capture=cap))
        s2 = e.stats()
        assert s2["prefix_hit_tokens"] == 192                      # the 3 full pages of the shared system prompt
        assert s2["tokens_prefill"] - s1["tokens_prefill"] == len(p2) - 192
        check_tokens_against(oracle, p2, o2[0], e.captured_logits(r2[0], cap))
        # next turn of chat 1: prompt + assistant reply + new user text; the reply's pages were cached as they filled
        p3 = np.concatenate([p1, np.array(o1[0], np.int32), weights.random_tokens(303, 30, d.vocab)])
        r3, o3 = e.generate([p3], hb.Sampling(max_tokens=6, capture=cap))
        s3 = e.stats()
        assert s3["prefix_hit_tokens"] - s2["prefix_hit_tokens"] == 256
        check_tokens_against(oracle, p3, o3[0], e.captured_logits(r3[0], cap))
        # an exact repeat leaves one page to compute (the last prompt token's logits are needed)
        r4, o4 = e.generate([p1], hb.Sampling(max_tokens=20))
        assert e.stats()["prefix_hit_tokens"] - s3["prefix_hit_tokens"] == 192 and len(o4[0]) == 20
        check_greedy(oracle, p1, o4[0])
        # two requests sharing the prefix in one step, while both hold references
        ra = e.submit(p1, hb.Sampling(max_tokens=5))
        rb = e.submit(p2, hb.Sampling(max_tokens=5))
        outs = {ra: [], rb: []}
        done = set()
        while len(done) < 2:
            e.step()
            for r in outs:
                t, fin = e.poll(r)
                outs[r] += t
                if fin:
                    done.add(r)
        check_greedy(oracle, p1, outs[ra])
        check_greedy(oracle, p2, outs[rb])
        st = e.stats()
        assert st["kv_pages_free"] == st["kv_pages_total"] and 0 < st["kv_pages_cached"] <= st["kv_pages_total"]
    # eviction: a 16-page pool, every request needs 6-7 pages and leaves cached pages behind
    with hb.Engine(hb.EngineConfig(max_seqs=2, max_ctx=512, max_batched_tokens=512, enable_prefix_cache=1)) as e:
        e.load_state_dict(d, sd)
        total = e.stats()["kv_pages_total"]
        assert total == 16
        for i in range(8):
            pr = weights.random_tokens(400 + i, 330 + i, d.vocab)
            _, o = e.generate([pr, pr[:100]], hb.Sampling(max_tokens=40))
            assert len(o[0]) == 40 and len(o[1]) == 40
            st = e.stats()
            assert st["kv_pages_free"] == total and st["running"] == 0
        assert e.stats()["prefix_hit_tokens"] == 0
        _, o = e.generate([pr], hb.Sampling(max_tokens=4))      # the most recent prompt is still cached
        assert e.stats()["prefix_hit_tokens"] == 320


def test_step_loop_thread_eos_cancel_and_errors():
    d = configs.tiny_llama(layers=2, head_dim=64, vocab=1000)
    sd = weights.llama_state_dict(d, 0, 0.02)
    prompt = weights.random_tokens(1, 48, d.vocab)
    with hb.Engine(hb.EngineConfig(max_seqs=4, max_ctx=256, max_batched_tokens=256)) as e:
        with pytest.raises(hb.HBError):
            e.submit(prompt, hb.Sampling())  # no model yet
        e.load_state_dict(d, sd)
        with pytest.raises(hb.HBError):
            e.submit([1000], hb.Sampling())  # token id out of range
        with pytest.raises(hb.HBError):
            e.submit(list(range(256)), hb.Sampling())  # prompt >= context_length
        e.start()
        r1 = e.submit(prompt, hb.Sampling(max_tokens=12, eos_token=303))  # golden greedy token is 303 -> stops at 1
        r2 = e.submit(prompt, hb.Sampling(max_tokens=200))
        assert e.wait(r1, 20000)
        toks, fin = [], 0
        while not fin:
            assert e.wait(r1, 20000)
            t, fin = e.poll(r1)
            toks += t
        assert toks == [303] and fin == 1
        e.cancel(r2)
        fin = 0
        while not fin:
            e.wait(r2, 20000)
            _, fin = e.poll(r2)
        assert fin == 2
        e.release(r1)
        e.release(r2)
        with pytest.raises(hb.HBError):
            e.poll(r1)
        # temperature sampling is reproducible per seed
        a = e.submit(prompt, hb.Sampling(max_tokens=8, temperature=0.8, seed=42))
        b = e.submit(prompt, hb.Sampling(max_tokens=8, temperature=0.8, seed=42))
        res = {}
        for r in (a, b):
            out, fin = [], 0
            while not fin:
                e.wait(r, 20000)
                t, fin = e.poll(r)
                out += t
            res[r] = out
        assert res[a] == res[b] and len(res[a]) == 8
        e.stop()
        st = e.stats()
        assert st["kv_pages_free"] == st["kv_pages_total"]


def test_memory_budget_contract():
    d = configs.tiny_llama(layers=2, head_dim=64, vocab=1000)
    cfg = hb.EngineConfig(max_seqs=4, max_ctx=256, max_batched_tokens=256)
    est = hb.engine.memory_estimate(d, cfg)
    with hb.Engine(cfg) as e:
        e.load_random(d, 1)
        st = e.stats()
        assert st["weights_bytes"] == est["weights"]
        assert st["kv_bytes"] <= est["kv"]
    tight = hb.EngineConfig(max_seqs=4, max_ctx=256, max_batched_tokens=256, memory_budget_bytes=est["weights"] + 1000)
    with hb.Engine(tight) as e:
        with pytest.raises(hb.HBError) as ei:
            e.load_random(d, 1)
        assert ei.value.code == -3  # HB_ERR_OOM: the scheduler's packing contract is enforced, not exceeded
    fit = est["weights"] + est["workspace"] + est["kv"] + (64 << 20)
    with hb.Engine(hb.EngineConfig(max_seqs=4, max_ctx=256, max_batched_tokens=256, memory_budget_bytes=fit)) as e:
        e.load_random(d, 1)
        st = e.stats()
        assert st["weights_bytes"] + st["kv_bytes"] + st["workspace_bytes"] <= fit


def test_bert_embed_matches_golden_and_oracle(golden_dir):
    g = np.load(os.path.join(golden_dir, "bert_tiny.npz"))
    d = configs.tiny_bert(layers=2, vocab=1000)
    sd = weights.bert_state_dict(d, int(g["seed"]), float(g["std"]))
    lens = g["lens"].tolist()
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    with hb.Engine(hb.EngineConfig(max_seqs=8, max_ctx=512, max_batched_tokens=700)) as e:
        e.load_state_dict(d, sd)
        out = e.embed_flat(g["tokens"], offs)   # 1013 tokens > 700: forces two engine batches
        assert np.abs(out - g["embeddings"]).max() <= 1e-2
        cos = (out * g["embeddings"]).sum(-1)
        assert cos.min() >= 0.9999
        # ragged + order independence: each sequence alone gives the same vector (bit-exact)
        for i in (0, 3, 4):
            solo = e.embed([g["tokens"][offs[i]:offs[i + 1]]])
            assert np.array_equal(solo[0], out[i])
        assert e.embed([]).shape == (0, d.hidden)
        with pytest.raises(hb.HBError):
            e.embed([list(range(513))])
        with pytest.raises(hb.HBError):
            e.submit([1, 2], hb.Sampling())
    seqs = [g["tokens"][offs[i]:offs[i + 1]] for i in range(len(lens))]
    assert np.abs(out - bert_embed(d, sd, seqs)).max() <= 1e-2


def test_runtime_lifecycle_and_openai_http_front():
    """runner.Runtime mirror end to end on the GPU: Start -> Warm -> URL() serves the three OpenAI routes -> Stop."""
    import json
    import urllib.request
    from helix_b200.runtime import B200Runtime, B200RuntimeParams

    d = configs.tiny_llama(layers=2, head_dim=64, vocab=1000)
    sd_rt = weights.llama_state_dict(d, 0, 0.02)
    rt = B200Runtime(B200RuntimeParams(model="tiny", desc=d, state_dict=sd_rt,
                                       args=["--max-num-seqs", "4", "--max-model-len", "256"]))
    assert rt.status() == ""
    rt.start()
    try:
        rt.warm("tiny")
        assert rt.status().startswith("running") and rt.runtime() == "vllm" and rt.list_models() == ["tiny"]
        base = rt.url()
        models = json.load(urllib.request.urlopen(base + "/v1/models"))
        assert models["data"][0]["id"] == "tiny"

        def post(path, body):
            req = urllib.request.Request(base + path, json.dumps(body).encode(), {"Content-Type": "application/json"})
            return urllib.request.urlopen(req, timeout=60)
        r = json.load(post("/v1/chat/completions", {"model": "tiny", "messages": [{"role": "user", "content": "hello"}],
                                                     "max_tokens": 6}))
        assert r["object"] == "chat.completion" and r["choices"][0]["finish_reason"] in ("stop", "length")
        assert r["usage"]["completion_tokens"] == 6 and r["usage"]["prompt_tokens"] > 0
        assert r["usage"]["total_tokens"] == r["usage"]["prompt_tokens"] + 6
        # what came over HTTP is the ORACLE's greedy continuation of the templated prompt, with the oracle's log-probabilities
        from helix_b200.server import ByteTokenizer
        from oracle import sampling_ref
        tk = ByteTokenizer()
        ids_in = tk.chat([{"role": "user", "content": "hello"}])
        o = LlamaOracle(d, sd_rt)
        want, rows = o.greedy(np.array(ids_in, np.int32), 6)
        assert r["choices"][0]["message"]["content"] == tk.decode([t for t in want if t != tk.EOS])
        rl = json.load(post("/v1/chat/completions", {"model": "tiny", "messages": [{"role": "user", "content": "hello"}],
                                                      "max_tokens": 6, "logprobs": True, "top_logprobs": 3, "n": 2, "seed": 3}))
        assert [c["message"]["content"] for c in rl["choices"]] == [r["choices"][0]["message"]["content"]] * 2  # greedy: both choices equal
        lp = rl["choices"][1]["logprobs"]["content"]
        for i, t in enumerate(want):
            _, ref_lp = sampling_ref.logprob_record(rows[i], t, 1)
            assert abs(lp[i]["logprob"] - ref_lp[0]) < 5e-2 and len(lp[i]["top_logprobs"]) == 3
            assert lp[i]["top_logprobs"][0]["logprob"] >= lp[i]["top_logprobs"][1]["logprob"] >= lp[i]["top_logprobs"][2]["logprob"]
        # `stop`: the byte tokenizer's output is arbitrary text; take a piece of it as the stop string of a rerun
        full = r["choices"][0]["message"]["content"]
        if len(full) >= 3:
            r2 = json.load(post("/v1/chat/completions", {"model": "tiny", "max_tokens": 6, "stop": [full[2:3]],
                                                          "messages": [{"role": "user", "content": "hello"}]}))
            assert r2["choices"][0]["message"]["content"] == full[:full.index(full[2:3])]
            assert r2["choices"][0]["finish_reason"] == "stop"
        lines = [l for l in post("/v1/chat/completions", {"model": "tiny", "stream": True, "max_tokens": 5,
                                                          "messages": [{"role": "user", "content": "hello"}]}).read().decode().split("\n\n") if l]
        assert lines[-1] == "data: [DONE]" and all(l.startswith("data: ") for l in lines)
        last = json.loads(lines[-2][6:])
        assert last["choices"][0]["finish_reason"]
        try:
            post("/v1/chat/completions", {"model": "wrong", "messages": []})
            assert False
        except urllib.error.HTTPError as e:
            assert e.code == 400
    finally:
        rt.stop()
    assert rt.status() == ""
    # encoder runtime (--task embed)
    b = configs.tiny_bert(layers=2, vocab=1000)
    sd_b = weights.bert_state_dict(b, 5, 0.05)
    rt = B200Runtime(B200RuntimeParams(model="tiny-embed", desc=b, state_dict=sd_b,
                                       args=["--task", "embed", "--max-model-len", "512"]))
    rt.start()
    try:
        rt.warm("tiny-embed")
        req = urllib.request.Request(rt.url() + "/v1/embeddings", json.dumps({"input": ["abc", "defgh"], "model": "tiny-embed"}).encode(),
                                     {"Content-Type": "application/json"})
        out = json.load(urllib.request.urlopen(req, timeout=60))
        assert len(out["data"]) == 2 and len(out["data"][0]["embedding"]) == b.hidden
        v = np.array(out["data"][1]["embedding"])
        assert abs(np.linalg.norm(v) - 1.0) < 1e-3
        # the vectors that came over HTTP are the oracle's embeddings of the tokenised inputs (all three input forms)
        from helix_b200.server import ByteTokenizer
        tk = ByteTokenizer()
        ref = bert_embed(b, sd_b, [np.array(tk.encode(t), np.int32) for t in ("abc", "defgh")])
        got = np.array([x["embedding"] for x in out["data"]], np.float32)
        assert np.abs(got - ref).max() <= 1e-2 and float((got * ref).sum(-1).min()) >= 0.9999
        req = urllib.request.Request(rt.url() + "/v1/embeddings", json.dumps({"input": [tk.encode("defgh")], "model": "tiny-embed"}).encode(),
                                     {"Content-Type": "application/json"})
        ids_form = json.load(urllib.request.urlopen(req, timeout=60))
        assert np.allclose(np.array(ids_form["data"][0]["embedding"], np.float32), got[1], atol=1e-6)
    finally:
        rt.stop()


def test_multi_model_pack_on_one_gpu():
    """BASELINE configs[3] in miniature: a decoder and an encoder engine co-resident on one device (one CUDA stream each,
    budgets honoured), driven concurrently from two host threads, produce exactly their solo results."""
    import threading
    d = configs.tiny_llama(layers=2, head_dim=64, vocab=1000)
    b = configs.tiny_bert(layers=2, vocab=1000)
    sd, sb = weights.llama_state_dict(d, 7, 0.05), weights.bert_state_dict(b, 5, 0.05)
    prompts = [weights.random_tokens(200 + i, n, d.vocab) for i, n in enumerate([40, 130, 7, 300])]
    seqs = [weights.random_tokens(300 + i, n, b.vocab) for i, n in enumerate([64, 512, 3, 200, 129])]
    cfg = dict(max_seqs=4, max_ctx=512, max_batched_tokens=1024)
    est_l = hb.engine.memory_estimate(d, hb.EngineConfig(**cfg))
    est_b = hb.engine.memory_estimate(b, hb.EngineConfig(**cfg))
    with hb.Engine(hb.EngineConfig(memory_budget_bytes=sum(est_l.values()) + (64 << 20), **cfg)) as el, \
            hb.Engine(hb.EngineConfig(memory_budget_bytes=sum(est_b.values()) + (64 << 20), **cfg)) as eb:
        el.load_state_dict(d, sd)
        eb.load_state_dict(b, sb)
        solo_gen = el.generate(prompts, hb.Sampling(max_tokens=12))[1]
        solo_emb = eb.embed(seqs)
        res = {}

        def chat():
            for _ in range(3):
                res["gen"] = el.generate(prompts, hb.Sampling(max_tokens=12))[1]

        def embed():
            for _ in range(6):
                res["emb"] = eb.embed(seqs)
        ts = [threading.Thread(target=chat), threading.Thread(target=embed)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert res["gen"] == solo_gen
        assert np.array_equal(res["emb"], solo_emb)
        sl, sbb = el.stats(), eb.stats()
        assert sl["weights_bytes"] + sl["kv_bytes"] + sl["workspace_bytes"] <= sl["budget_bytes"]
        assert sbb["weights_bytes"] + sbb["workspace_bytes"] <= sbb["budget_bytes"]


def test_safetensors_checkpoint_load_is_bit_identical(tmp_path):
    from helix_b200 import weights_io
    d = configs.tiny_llama(layers=2, head_dim=64, vocab=1000)
    sd = weights.llama_state_dict(d, 0, 0.05)
    path = tmp_path / "model.safetensors"
    weights_io.write_safetensors(path, sd)
    prompt = weights.random_tokens(1, 48, d.vocab)
    cfg = hb.EngineConfig(max_seqs=4, max_ctx=256, max_batched_tokens=256)
    with hb.Engine(cfg) as a, hb.Engine(cfg) as b:
        a.load_state_dict(d, sd)
        weights_io.load_safetensors(b, d, path)
        ra, oa = a.generate([prompt], hb.Sampling(max_tokens=6, capture=CAPTURE_STEP_LOGITS))
        rb, ob = b.generate([prompt], hb.Sampling(max_tokens=6, capture=CAPTURE_STEP_LOGITS))
        assert oa == ob
        assert np.array_equal(a.captured_logits(ra[0], CAPTURE_STEP_LOGITS), b.captured_logits(rb[0], CAPTURE_STEP_LOGITS))


def test_random_arrivals_cancellations_never_leak_pages_and_stay_deterministic():
    """Scheduler bookkeeping under churn (bit-exact integer work): requests arrive while others decode, a third are
    cancelled at random points, the queue outruns max_seqs; afterwards every KV page is back, and every request that
    ran to completion produced exactly the tokens it produces alone."""
    d = configs.tiny_llama(layers=2, head_dim=64, vocab=1000)
    sd = weights.llama_state_dict(d, 9, 0.05)
    rng = np.random.default_rng(0)
    n = 40
    prompts = [weights.random_tokens(400 + i, int(rng.integers(1, 300)), d.vocab) for i in range(n)]
    max_new = [int(rng.integers(1, 24)) for _ in range(n)]
    with hb.Engine(hb.EngineConfig(max_seqs=6, max_ctx=384, max_batched_tokens=512, use_cuda_graphs=1)) as e:
        e.load_state_dict(d, sd)
        solo = {}
        for i in (0, 7, 13, 21, 39):
            solo[i] = e.generate([prompts[i]], hb.Sampling(max_tokens=max_new[i]))[1][0]
        total_pages = e.stats()["kv_pages_total"]
        rids, outs, done, cancelled = {}, {}, set(), set()
        submitted, steps = 0, 0
        while len(done) < n:
            while submitted < n and rng.random() < 0.6:   # bursty arrivals
                rids[submitted] = e.submit(prompts[submitted], hb.Sampling(max_tokens=max_new[submitted]))
                outs[submitted] = []
                submitted += 1
            e.step()
            steps += 1
            st = e.stats()
            assert st["running"] <= 6 and 0 <= st["kv_pages_free"] <= total_pages
            for i, r in list(rids.items()):
                if i in done:
                    continue
                if i not in solo and i % 3 == 0 and i not in cancelled and rng.random() < 0.3:
                    e.cancel(r)
                    cancelled.add(i)
                t, fin = e.poll(r)
                outs[i] += t
                if fin:
                    done.add(i)
            assert steps < 5000
        for _ in range(3):
            e.step()   # retire anything cancelled on the last pass
        st = e.stats()
        assert st["kv_pages_free"] == total_pages and st["running"] == 0 and st["waiting"] == 0
        for i, want in solo.items():
            assert outs[i] == want
        for i in range(n):
            if i not in cancelled:
                assert len(outs[i]) == max_new[i]


def test_churn_with_chunked_prefill_prefix_cache_and_sampling_filters():
    """Everything at once: a 96-token step budget (most prompts are chunked), the prefix cache on with prompts drawn
    from three shared prefixes (hits, shared pages held by several sequences, evictions in a small pool), random
    cancellations (also of half-prefilled prompts), a mix of greedy / temperature / top-k / top-p requests.
    Afterwards no page is leaked, every uncancelled request has its full length and greedy ones agree with the oracle;
    a second pass over the same requests (now mostly prefix hits) must satisfy the same."""
    d = configs.tiny_llama(layers=2, head_dim=64, vocab=1000)
    sd = weights.llama_state_dict(d, 17, 0.05)
    oracle = LlamaOracle(d, sd)
    rng = np.random.default_rng(3)
    prefixes = [weights.random_tokens(700 + i, 130 + 40 * i, d.vocab) for i in range(3)]
    n = 30
    prompts, samp = [], []
    for i in range(n):
        tail = weights.random_tokens(800 + i, int(rng.integers(1, 120)), d.vocab)
        prompts.append(np.concatenate([prefixes[i % 3], tail]) if i % 4 else tail)
        kind = i % 4
        m = int(rng.integers(2, 20))
        samp.append([hb.Sampling(max_tokens=m), hb.Sampling(max_tokens=m, temperature=0.8, seed=i),
                     hb.Sampling(max_tokens=m, temperature=1.1, seed=i, top_k=5),
                     hb.Sampling(max_tokens=m, temperature=0.9, seed=i, top_p=0.6)][kind])

    def run(e, cancel_some):
        rids, outs, done, cancelled = {}, {}, set(), set()
        submitted, steps = 0, 0
        total = e.stats()["kv_pages_total"]
        while len(done) < n:
            while submitted < n and rng.random() < 0.5:
                rids[submitted] = e.submit(prompts[submitted], samp[submitted])
                outs[submitted] = []
                submitted += 1
            e.step()
            steps += 1
            st = e.stats()
            assert st["running"] <= 4 and 0 <= st["kv_pages_free"] <= total
            for i, r in list(rids.items()):
                if i in done:
                    continue
                if cancel_some and i % 5 == 2 and i not in cancelled and rng.random() < 0.25:
                    e.cancel(r)
                    cancelled.add(i)
                t, fin = e.poll(r)
                outs[i] += t
                if fin:
                    done.add(i)
            assert steps < 5000
        for _ in range(3):
            e.step()
        st = e.stats()
        assert st["kv_pages_free"] == total and st["running"] == 0 and st["waiting"] == 0
        for r in rids.values():
            e.release(r)
        return outs, cancelled

    with hb.Engine(hb.EngineConfig(max_seqs=4, max_ctx=512, max_batched_tokens=96, use_cuda_graphs=1,
                                   enable_prefix_cache=1)) as e:
        e.load_state_dict(d, sd)
        outs1, cancelled = run(e, True)
        hits1 = e.stats()["prefix_hit_tokens"]
        outs2, _ = run(e, False)
        assert e.stats()["prefix_hit_tokens"] > hits1 > 0
    for i in range(n):
        assert len(outs2[i]) == samp[i].max_tokens
        if i not in cancelled:
            assert len(outs1[i]) == samp[i].max_tokens
        if i % 4 == 0:
            check_greedy(oracle, prompts[i], outs2[i])
            if i not in cancelled:
                check_greedy(oracle, prompts[i], outs1[i])
