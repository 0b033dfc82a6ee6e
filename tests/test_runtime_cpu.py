"""CPU: host-side Runtime mirror — vLLM-arg contract (bit-exact integers) and OpenAI wire framing (stub engine)."""
import json

import pytest

from helix_b200 import runtime as R
from helix_b200.server import ByteTokenizer, OpenAIServer, chat_chunk
from oracle import scheduler_ref as S

GB = 1024 ** 3


def test_args_from_the_scheduler_round_trip():
    # what RunnerController.substituteVLLMArgsPlaceholders emits (scheduler/runner.go:1222-1259) ...
    args = S.substitute_vllm_args(["--max-model-len", "8192", "--max-num-seqs", "64"], 80 * GB, 8 * GB)
    p = R.parse_vllm_args(["--host", "127.0.0.1", "--port", "1234"] + args)
    assert p.gpu_memory_utilization == 0.10 and p.max_num_seqs == 64 and p.max_model_len == 8192 and not p.task_embed
    assert R.parse_vllm_args(["--task", "embed", "--trust-remote-code"]).task_embed
    assert R.parse_vllm_args([]).max_num_seqs == 256  # types/memory.go:11
    assert R.parse_vllm_args([]).enable_prefix_caching and not R.parse_vllm_args(["--no-enable-prefix-caching"]).enable_prefix_caching
    assert R.parse_vllm_args(["--max-num-batched-tokens", "8192", "--enable-prefix-caching"]).max_num_batched_tokens == 8192
    # ... and the budget the engine derives: exact bytes when the slot carries them, ratio*perGPU otherwise
    assert R.memory_budget(8 * GB, 80 * GB, 0.10) == 8 * GB
    assert R.memory_budget(0, 80 * GB, 0.10) == 8 * GB
    assert R.memory_budget(0, 0, None) == 0


class StubEngine:
    """Scripted engine: yields the given token groups, one group per poll."""

    class _D:
        vocab = 260
        hidden = 4

    class _C:
        max_ctx = 64

    def __init__(self, groups):
        self.groups, self.i, self.desc, self.cfg = groups, 0, self._D(), self._C()
        self.cancelled = self.released = False

    def submit(self, ids, sp):
        self.prompt, self.sp = list(ids), sp
        return 7

    def wait(self, rid, ms):
        return True

    def poll(self, rid):
        g = self.groups[self.i]
        self.i += 1
        return g, (1 if self.i == len(self.groups) else 0)

    def cancel(self, rid):
        self.cancelled = True

    def release(self, rid):
        self.released = True

    def embed(self, seqs):
        return [[float(len(s)), 0.0, 0.0, 1.0] for s in seqs]


class StubRuntime:
    def __init__(self, eng):
        self.engine = eng
        self.p = R.B200RuntimeParams(model="m")

    def list_models(self):
        return ["m"]

    def status(self):
        return "running"


def test_chat_sse_framing_and_finish_reason():
    tok = ByteTokenizer()
    hi = [b + tok.OFFSET for b in b"hi"]
    eng = StubEngine([hi[:1], hi[1:], []])
    srv = OpenAIServer(StubRuntime(eng), tok)
    chunks = list(srv.chat_stream({"model": "m", "messages": [{"role": "user", "content": "x"}], "max_tokens": 5, "stream": True}))
    assert chunks[0]["choices"][0]["delta"] == {"role": "assistant", "content": ""}
    assert "".join(c["choices"][0]["delta"].get("content", "") for c in chunks) == "hi"
    assert [c["choices"][0]["finish_reason"] for c in chunks[:-1]] == [None] * (len(chunks) - 1)
    assert chunks[-1]["choices"][0]["finish_reason"] == "stop"      # non-empty: closes the control-plane stream
    assert all(c["object"] == "chat.completion.chunk" and c["id"] == chunks[0]["id"] for c in chunks)
    assert eng.released and eng.prompt[0] == tok.BOS and eng.sp.max_tokens == 5
    json.dumps(chunks)  # serialisable
    with pytest.raises(ValueError):
        list(srv.chat_stream({"model": "other", "messages": []}))
    eng2 = StubEngine([hi, hi[:1]])
    out = OpenAIServer(StubRuntime(eng2), tok).chat({"messages": [{"role": "user", "content": "x"}], "max_tokens": 3})
    assert out["choices"][0]["message"]["content"] == "hih" and out["choices"][0]["finish_reason"] == "length"


def test_embeddings_accepts_all_reference_input_forms():
    srv = OpenAIServer(StubRuntime(StubEngine([])), ByteTokenizer())
    for inp, n in (("abc", 1), (["a", "bcd"], 2), ([[1, 2, 3], [4]], 2), ([5, 6], 1)):
        r = srv.embeddings({"input": inp, "model": "m"})
        assert len(r["data"]) == n and r["data"][0]["object"] == "embedding" and len(r["data"][0]["embedding"]) == 4
        assert r["usage"]["prompt_tokens"] > 0
    assert srv.models()["data"][0]["id"] == "m"
    assert chat_chunk("i", "m", 1, {}, "stop")["choices"][0]["finish_reason"] == "stop"


def test_embedding_micro_batcher_coalesces_concurrent_single_chunk_requests():
    """The RAG caller's pattern (1 chunk per request, 10 workers): requests landing inside the window share one hb_embed."""
    import threading
    import time
    from helix_b200.server import EmbedBatcher

    class CountingEngine(StubEngine):
        def __init__(self):
            super().__init__([])
            self.calls = []

        def embed(self, seqs):
            self.calls.append(len(seqs))
            time.sleep(0.01)
            if any(len(s) == 0 for s in seqs):
                raise ValueError("empty sequence")
            return [[float(s[0]), 0.0, 0.0, 1.0] for s in seqs]

    eng = CountingEngine()
    b = EmbedBatcher(eng, window_s=0.05, max_seqs=64)
    out = {}

    def worker(i):
        out[i] = b.embed([[i, i + 1]])
    ts = [threading.Thread(target=worker, args=(i,)) for i in range(10)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert all(out[i][0][0] == float(i) for i in range(10))       # every caller got ITS vector back
    assert sum(eng.calls) == 10 and len(eng.calls) <= 3            # coalesced (normally a single call)
    # a failing request is isolated: its neighbours still succeed
    res = {}

    def worker2(i, seq):
        try:
            res[i] = b.embed([seq])
        except ValueError as e:
            res[i] = e
    ts = [threading.Thread(target=worker2, args=(0, [])), threading.Thread(target=worker2, args=(1, [5]))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert isinstance(res[0], ValueError) and res[1][0][0] == 5.0
    b.close()


def test_hf_tokenizer_wrapper_and_llama3_chat_template(tmp_path):
    """Row F2: the `tokenizers`-backed wrapper (no real Llama-3 vocabulary exists offline, so a tiny BPE with the same
    special tokens is trained here): chat() must produce exactly the Llama-3 template, token ids bit-exact vs the library."""
    from tokenizers import Tokenizer, models, pre_tokenizers, trainers, decoders
    from helix_b200.server import HFTokenizer
    specials = ["<|begin_of_text|>", "<|start_header_id|>", "<|end_header_id|>", "<|eot_id|>"]
    tk = Tokenizer(models.BPE())
    tk.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
    tk.decoder = decoders.ByteLevel()
    tk.train_from_iterator(["hello world, say the word warm", "system user assistant"] * 20,
                           trainers.BpeTrainer(vocab_size=300, special_tokens=specials,
                                               initial_alphabet=pre_tokenizers.ByteLevel.alphabet()))
    path = str(tmp_path / "tokenizer.json")
    tk.save(path)
    w = HFTokenizer(path)
    msgs = [{"role": "system", "content": "be brief"}, {"role": "user", "content": "Say the word 'warm'."}]
    ids = w.chat(msgs)
    want = ("<|begin_of_text|><|start_header_id|>system<|end_header_id|>\n\nbe brief<|eot_id|>"
            "<|start_header_id|>user<|end_header_id|>\n\nSay the word 'warm'.<|eot_id|>"
            "<|start_header_id|>assistant<|end_header_id|>\n\n")
    assert ids == tk.encode(want, add_special_tokens=False).ids          # bit-exact ids vs the library
    assert ids[0] == tk.token_to_id("<|begin_of_text|>") and w.EOS == tk.token_to_id("<|eot_id|>")
    assert tk.decode(ids, skip_special_tokens=False) == want
    assert w.decode(w.encode("hello world")) == "hello world"


def test_stop_matcher_streaming_semantics():
    """OpenAI `stop`: cut at the first occurrence, never emit the stop text, hold back text that may still become one."""
    from helix_b200.server import StopMatcher
    m = StopMatcher(None)
    assert m.feed("abc") == "abc" and m.flush() == "" and not m.hit
    m = StopMatcher("END")
    out = [m.feed(t) for t in ["hello E", "N", "x EN", "D tail"]]
    assert out == ["hello ", "", "ENx ", ""] and m.hit and m.flush() == ""
    assert m.feed("more") == ""
    m = StopMatcher(["\n\n", "###"])
    assert m.feed("a\n") == "a" and m.feed("b#") == "\nb" and m.feed("#") == "" and m.flush() == "##" and not m.hit
    m = StopMatcher(["\n\n", "###"])
    assert "".join(m.feed(c) for c in "line one\n\nline two") == "line one" and m.hit
    # every split of a text yields the same visible output
    text, stops = "alpha <|eot|> beta <|eot|>", ["<|eot|>"]
    for cut in range(len(text) + 1):
        m = StopMatcher(stops)
        got = m.feed(text[:cut]) + m.feed(text[cut:]) + m.flush()
        assert got == "alpha ", (cut, got)


def test_memory_estimation_endpoint_shape_and_numbers():
    """POST /memory-estimate for this runtime (api/pkg/types/memory.go:16-50): closed form, one `single_gpu` plan whose
    total is what the engine will actually allocate and what the allocator packs with."""
    r = R.memory_estimation({"model_name": "meta-llama/Meta-Llama-3-8B-Instruct", "context_length": 2048,
                             "batch_size": 16384, "num_parallel": 32}, runner_id="r1")
    assert r["success"] and r["runner_id"] == "r1" and r["architecture"] == "llama" and r["block_count"] == 32
    (c,) = r["configurations"]
    assert c["name"] == "single_gpu" and c["gpu_count"] == 1 and c["fully_loaded"] and c["total_layers"] == 33
    assert abs(c["weights_memory"] - 16.06e9) < 0.02e9 and c["kv_cache"] == 32 * 2048 * 131072
    assert c["vram_required"] == c["total_memory"] == c["weights_memory"] + c["kv_cache"] + c["graph_memory"]
    assert c["gpu_sizes"] == [c["total_memory"]]
    # the scheduler packs with this number: two such instances plus the 1B model fit a 180 GB B200, a fourth 8B does not
    from oracle import scheduler_ref as S
    need = c["total_memory"]
    small = R.memory_estimation({"model_name": "meta-llama/Llama-3.2-1B-Instruct", "context_length": 2048, "num_parallel": 32})
    slots = [{"id": i, "model": f"m{i}", "runtime": "vllm", "gpus": [0], "memory": need, "stale": False, "last_activity": 0}
             for i in range(2)]
    gpus = [(0, 180 * 10 ** 9)]
    assert S.plan_allocation(gpus, slots, {"model": "s", "runtime": "vllm", "memory": small["configurations"][0]["total_memory"]})
    many = slots + [dict(slots[0], id=9, model="m9")] * 6
    assert S.plan_allocation(gpus, many, {"model": "x", "runtime": "vllm", "memory": need}) is None
    bad = R.memory_estimation({"model_name": "unknown/model"})
    assert not bad["success"] and "not served" in bad["error"] and bad["configurations"] == []
    emb = R.memory_estimation({"model_name": "BAAI/bge-base-en-v1.5", "context_length": 512, "num_parallel": 64})
    assert emb["success"] and emb["architecture"] == "bert" and emb["configurations"][0]["kv_cache"] == 0


class _StubEngine:
    """Records what the runtime asks of the engine; `fail_load` makes the weight load raise."""
    instances = []

    def __init__(self, cfg, fail_load=False):
        self.cfg, self.fail_load, self.closed, self.started, self.loaded = cfg, fail_load, False, False, None
        self.desc = None
        _StubEngine.instances.append(self)

    def load_random(self, desc, seed):
        if self.fail_load:
            raise R.HBError(-3, "weights exceed the memory budget")
        self.loaded, self.desc = ("random", seed), desc

    def load_state_dict(self, desc, sd):
        self.loaded, self.desc = ("state_dict", len(sd)), desc

    def start(self):
        self.started = True

    def close(self):
        self.closed = True

    def stats(self):
        return {"cuda_error": 0, "kv_pages_free": 7, "kv_pages_total": 8, "running": 0, "waiting": 0}

    def embed(self, seqs):
        return [[0.0]] * len(seqs)


def test_runtime_lifecycle_against_a_stub_engine():
    """Runtime.Start/Stop/Status/URL/ListModels semantics (api/pkg/runner/slot.go:46-57,113-140) and the engine
    configuration derived from CreateRunnerSlotAttributes + the scheduler's vLLM-style args — no GPU involved."""
    _StubEngine.instances.clear()
    GB = 1024 ** 3
    p = R.B200RuntimeParams(model="meta-llama/Meta-Llama-3-8B-Instruct", gpu_index=3, model_memory_requirement=40 * GB,
                            per_gpu_memory=180 * GB, context_length=4096, serve_http=False, seed=5,
                            args=["--gpu-memory-utilization", "0.22", "--max-num-seqs", "64", "--max-model-len", "8192",
                                  "--max-num-batched-tokens", "8192"], engine_factory=_StubEngine)
    rt = R.B200Runtime(p)
    assert rt.status() == "" and rt.url() == "" and rt.runtime() == "vllm" and rt.list_models() == [p.model]
    rt.start()
    (e,) = _StubEngine.instances
    c = e.cfg
    assert (c.device, c.memory_budget_bytes, c.max_seqs, c.max_ctx, c.max_batched_tokens) == (3, 40 * GB, 64, 8192, 8192)
    assert c.enable_prefix_cache == 1 and c.use_cuda_graphs == 1        # exact bytes beat the 2-decimal ratio flag
    assert e.loaded == ("random", 5) and e.started and e.desc.layers == 32
    assert rt.status().startswith("running") and "kv_pages_free=7/8" in rt.status()
    assert "helix-b200" in rt.command_line() and "--max-num-seqs 64" in rt.command_line()
    seen = []
    rt.pull_model(p.model, seen.append)
    assert seen and seen[-1]["status"] == "success"
    rt.stop()
    assert e.closed and rt.status() == "" and rt.engine is None
    # the ratio flag is the fallback when the attributes carry no byte count
    _StubEngine.instances.clear()
    rt = R.B200Runtime(R.B200RuntimeParams(model="BAAI/bge-base-en-v1.5", per_gpu_memory=100 * GB, serve_http=False,
                                           args=["--gpu-memory-utilization", "0.05", "--task", "embed"],
                                           engine_factory=_StubEngine))
    rt.start()
    (e,) = _StubEngine.instances
    assert e.cfg.memory_budget_bytes == 5 * GB and e.cfg.enable_prefix_cache == 0 and not e.started and rt.is_embed
    rt.warm("BAAI/bge-base-en-v1.5")
    rt.stop()
    # a failing load leaves nothing behind (Slot.Create's deferred Stop would otherwise leak device memory)
    _StubEngine.instances.clear()
    rt = R.B200Runtime(R.B200RuntimeParams(model="meta-llama/Llama-3.2-1B-Instruct", serve_http=False,
                                           engine_factory=lambda cfg: _StubEngine(cfg, fail_load=True)))
    with pytest.raises(R.HBError):
        rt.start()
    assert _StubEngine.instances[0].closed and rt.engine is None and rt.status() == ""
    with pytest.raises(R.HBError):
        R.B200Runtime(R.B200RuntimeParams(model="nobody/unknown", serve_http=False, engine_factory=_StubEngine)).start()
