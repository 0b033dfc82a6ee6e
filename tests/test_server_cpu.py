"""CPU: the OpenAI-compatible front (helix_b200/server.py) over real sockets, with a scripted stand-in for the engine —
the wire behaviour the reference's runner depends on: SSE framing, a final chunk with a non-empty
finish_reason, `data: [DONE]`, usage accounting, `stop`, request validation, embedding input forms and coalescing."""
import json
import threading
import types
import urllib.error
import urllib.request

import numpy as np
import pytest

from helix_b200.engine import HBError
from helix_b200.server import ByteTokenizer, OpenAIServer


class FakeEngine:
    """Generates a fixed text, a few tokens per poll; records cancel/release calls."""

    def __init__(self, text="Hello world!\n\nSecond paragraph."):
        self.tok = ByteTokenizer()
        self.script = [b + ByteTokenizer.OFFSET for b in text.encode()]
        self.desc = types.SimpleNamespace(vocab=1000, hidden=8)
        self.cfg = types.SimpleNamespace(max_ctx=64)
        self.reqs, self.cancelled, self.released, self.embed_calls = {}, [], [], []
        self.lock = threading.Lock()
        self.next_id = 1

    def submit(self, ids, sp):
        with self.lock:
            rid = self.next_id
            self.next_id += 1
            self.reqs[rid] = {"sp": sp, "pos": 0, "n_prompt": len(ids), "cancelled": False}
        return rid

    def wait(self, rid, timeout_ms):
        return True

    def poll(self, rid):
        r = self.reqs[rid]
        if r["cancelled"]:
            return [], 2
        n = min(3, r["sp"].max_tokens - r["pos"], len(self.script) - r["pos"])
        toks = self.script[r["pos"]:r["pos"] + n]
        r["pos"] += n
        fin = 1 if (r["pos"] >= r["sp"].max_tokens or r["pos"] >= len(self.script)) else 0
        return toks, fin

    def logprobs(self, rid, first_row, max_rows):
        r = self.reqs[rid]
        w = r["sp"].logprobs
        rows = self.script[first_row:min(r["pos"], first_row + max_rows)]
        ids = np.array([[t] + [t + 1 + j for j in range(w - 1)] for t in rows], np.int32).reshape(len(rows), w)
        lps = np.array([[-0.5] + [-1.0 - j for j in range(w - 1)] for _ in rows], np.float32).reshape(len(rows), w)
        return ids, lps

    def cancel(self, rid):
        self.reqs[rid]["cancelled"] = True
        self.cancelled.append(rid)

    def release(self, rid):
        self.released.append(rid)

    def embed(self, seqs):
        self.embed_calls.append(len(seqs))
        for s in seqs:
            if len(s) == 0:
                raise HBError(-1, "empty sequence")
        return np.stack([np.full(8, float(len(s)), np.float32) for s in seqs])


class FakeRuntime:
    def __init__(self):
        self.engine = FakeEngine()
        self.p = types.SimpleNamespace(model="tiny")

    def list_models(self):
        return ["tiny"]

    def status(self):
        return "running"


@pytest.fixture()
def front():
    rt = FakeRuntime()
    srv = OpenAIServer(rt)
    base = srv.start()
    yield rt, base
    srv.stop()


def post(base, path, body, raw=False):
    req = urllib.request.Request(base + path, json.dumps(body).encode(), {"Content-Type": "application/json"})
    r = urllib.request.urlopen(req, timeout=20)
    return r.read().decode() if raw else json.load(r)


def test_models_and_health(front):
    rt, base = front
    assert json.load(urllib.request.urlopen(base + "/v1/models"))["data"][0]["id"] == "tiny"   # slot.go:606-624
    assert json.load(urllib.request.urlopen(base + "/healthz"))["status"] == "running"


def test_chat_json_usage_and_length(front):
    rt, base = front
    r = post(base, "/v1/chat/completions", {"model": "tiny", "messages": [{"role": "user", "content": "hi"}], "max_tokens": 5})
    assert r["object"] == "chat.completion" and r["choices"][0]["message"]["content"] == "Hello"
    assert r["choices"][0]["finish_reason"] == "length"
    n_prompt = len(ByteTokenizer().chat([{"role": "user", "content": "hi"}]))
    assert r["usage"] == {"prompt_tokens": n_prompt, "completion_tokens": 5, "total_tokens": n_prompt + 5}
    assert rt.engine.released == [1] and rt.engine.cancelled == []


def test_chat_stream_sse_framing(front):
    rt, base = front
    raw = post(base, "/v1/chat/completions", {"model": "tiny", "stream": True, "max_tokens": 200,
                                              "messages": [{"role": "user", "content": "hi"}]}, raw=True)
    events = [e for e in raw.split("\n\n") if e]
    assert all(e.startswith("data: ") for e in events) and events[-1] == "data: [DONE]"
    chunks = [json.loads(e[6:]) for e in events[:-1]]
    assert chunks[0]["choices"][0]["delta"] == {"role": "assistant", "content": ""}
    assert all(c["object"] == "chat.completion.chunk" and c["id"] == chunks[0]["id"] for c in chunks)
    assert [c["choices"][0]["finish_reason"] for c in chunks[:-1]] == [None] * (len(chunks) - 1)
    assert chunks[-1]["choices"][0]["finish_reason"] == "stop"       # closes the control-plane stream (helix_openai_client.go:197)
    text = "".join(c["choices"][0]["delta"].get("content", "") for c in chunks)
    assert text == "Hello world!\n\nSecond paragraph."
    assert chunks[-1]["usage"]["completion_tokens"] == len(text.encode())


def test_stop_sequence_cuts_and_cancels(front):
    rt, base = front
    r = post(base, "/v1/chat/completions", {"model": "tiny", "max_tokens": 200, "stop": ["\n\n"],
                                            "messages": [{"role": "user", "content": "hi"}]})
    assert r["choices"][0]["message"]["content"] == "Hello world!" and r["choices"][0]["finish_reason"] == "stop"
    assert rt.engine.cancelled == [1] and rt.engine.released == [1]   # the sequence's KV pages are freed, the record released


def test_request_validation(front):
    rt, base = front
    with pytest.raises(urllib.error.HTTPError) as e:
        post(base, "/v1/chat/completions", {"model": "other", "messages": []})
    assert e.value.code == 400 and "model mismatch" in json.load(e.value)["error"]["message"]   # openai_chat_handlers.go:44-50
    with pytest.raises(urllib.error.HTTPError) as e:
        post(base, "/v1/embeddings", {"model": "tiny", "input": 7})
    assert e.value.code == 400
    with pytest.raises(urllib.error.HTTPError) as e:
        post(base, "/v1/nothing", {})
    assert e.value.code == 404


def test_embedding_input_forms_and_coalescing(front):
    rt, base = front
    one = post(base, "/v1/embeddings", {"model": "tiny", "input": "abc"})                 # string
    many = post(base, "/v1/embeddings", {"model": "tiny", "input": ["a", "bcd"]})         # []string
    ids = post(base, "/v1/embeddings", {"model": "tiny", "input": [[5, 6, 7], [8]]})      # [][]int (types/types.go:2707-2730)
    flat = post(base, "/v1/embeddings", {"model": "tiny", "input": [5, 6, 7, 8]})         # []int = one sequence
    assert [d["index"] for d in many["data"]] == [0, 1] and one["data"][0]["embedding"][0] == 4.0   # BOS + 3 bytes
    assert [d["embedding"][0] for d in ids["data"]] == [3.0, 1.0] and flat["data"][0]["embedding"][0] == 4.0
    assert ids["usage"] == {"prompt_tokens": 4, "total_tokens": 4}
    # the RAG caller sends one chunk per request from 10 workers: the batcher turns them into few engine calls
    before = len(rt.engine.embed_calls)
    out = [None] * 10

    def worker(i):
        out[i] = post(base, "/v1/embeddings", {"model": "tiny", "input": "x" * (i + 1)})

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(10)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert [o["data"][0]["embedding"][0] for o in out] == [float(i + 2) for i in range(10)]
    calls = rt.engine.embed_calls[before:]
    assert sum(calls) == 10 and len(calls) <= 10


def test_streamed_text_does_not_depend_on_poll_boundaries(front):
    """Multi-byte characters whose bytes straddle two polls (the stub hands out 3 tokens per poll) must come out whole,
    and identically in streaming and non-streaming mode."""
    rt, base = front
    text = "naïve café — 日本語 ✓ end"
    rt.engine.script = [b + ByteTokenizer.OFFSET for b in text.encode()]
    body = {"model": "tiny", "max_tokens": 500, "messages": [{"role": "user", "content": "hi"}]}
    assert post(base, "/v1/chat/completions", body)["choices"][0]["message"]["content"] == text
    raw = post(base, "/v1/chat/completions", dict(body, stream=True), raw=True)
    chunks = [json.loads(e[6:]) for e in raw.split("\n\n") if e and e != "data: [DONE]"]
    pieces = [c["choices"][0]["delta"].get("content", "") for c in chunks]
    assert "".join(pieces) == text and all("�" not in p for p in pieces)
    # a stop string that itself straddles polls and multi-byte characters
    r = post(base, "/v1/chat/completions", dict(body, stop=["— 日"]))
    assert r["choices"][0]["message"]["content"] == "naïve café " and r["choices"][0]["finish_reason"] == "stop"
    # max_tokens cutting a character in half: the dangling bytes surface as one replacement char at the end, never earlier
    cut = len("naï".encode()) - 1
    r = post(base, "/v1/chat/completions", dict(body, max_tokens=cut))
    assert r["choices"][0]["message"]["content"] == "na�" and r["choices"][0]["finish_reason"] == "length"


def test_stream_decoder_and_stop_matcher_properties():
    """Property checks (hypothesis): any grouping of the token stream into polls yields the same visible text as decoding
    everything at once; with stop strings, the visible text is the full text cut at the first stop occurrence."""
    from hypothesis import given, settings, strategies as st
    from helix_b200.server import StopMatcher, StreamDecoder
    tok = ByteTokenizer()
    alphabet = st.sampled_from(list("ab é日✓\n#"))

    @settings(max_examples=150, deadline=None)
    @given(st.lists(alphabet, max_size=30), st.lists(st.integers(1, 5), min_size=1, max_size=40),
           st.lists(st.text(alphabet=list("ab\n#é"), min_size=1, max_size=3), max_size=2))
    def prop(chars, cuts, stops):
        text = "".join(chars)
        ids = [b + ByteTokenizer.OFFSET for b in text.encode()]
        dec, sm, out, i, k = StreamDecoder(tok), StopMatcher(stops), "", 0, 0
        while i < len(ids) and not sm.hit:
            n = cuts[k % len(cuts)]
            k += 1
            piece = ids[i:i + n]
            i += n
            out += sm.feed(dec.feed(piece, final=i >= len(ids)))
        out += sm.flush()
        first = min((text.find(s) for s in stops if s in text), default=-1)
        assert out == (text if first < 0 else text[:first])

    prop()


def test_streaming_decode_of_arbitrary_bytes_matches_whole_decode():
    """A random-init model emits invalid UTF-8 all the time: the incremental decoder must still show exactly what decoding
    the whole sequence shows, whatever the grouping (and must not hold text back forever behind an invalid byte)."""
    import random
    from helix_b200.server import StreamDecoder
    tok, rnd = ByteTokenizer(), random.Random(7)
    for trial in range(600):
        ids = [rnd.randrange(256) + ByteTokenizer.OFFSET for _ in range(rnd.randrange(0, 80))]
        dec, out, i, held = StreamDecoder(tok), "", 0, 0
        while i < len(ids):
            k = rnd.randrange(1, 7)
            i += k
            piece = dec.feed(ids[i - k:i], final=i >= len(ids))
            held = 0 if piece else held + k
            assert held <= 16   # bounded hold-back
            out += piece
        assert out == tok.decode(ids)


def test_native_float_array_encoder_round_trips_float32():
    """hb_json_f32_array (the embedding response's float encoder): valid JSON, shortest text that parses back to the
    same float32 — what Go's encoding/json produces for []float32 — null for non-finite values, size query honoured."""
    import ctypes
    from helix_b200 import _lib
    from helix_b200.server import native_f32_json
    fmt, L = native_f32_json(), _lib.lib()
    assert fmt is not None
    rng = np.random.default_rng(3)
    for scale in (1.0, 1e-6, 1e12, 3e-39):
        v = (rng.standard_normal(1000) * scale).astype(np.float32)
        text = fmt(v)
        assert np.array_equal(np.array(json.loads(text), dtype=np.float32), v)
        assert len(text) < len(json.dumps(v.tolist()))          # shortest float32 text, not the double's 17 digits
    special = np.array([0.0, -0.0, 1.0, 0.1, 16777216.0, 3.4028235e38, 1e-45, np.nan, np.inf, -np.inf], np.float32)
    back = json.loads(fmt(special))
    assert back[7:] == [None, None, None]
    assert np.array_equal(np.array(back[:7], dtype=np.float32), special[:7]) and fmt(special).startswith(b"[0,-0,1,0.1,16777216,")
    assert fmt(np.zeros(0, np.float32)) == b"[]"
    v = np.arange(8, dtype=np.float32)
    buf = ctypes.create_string_buffer(4)
    need = L.hb_json_f32_array(v.ctypes.data, v.size, buf, 4)    # too small: reports the size, writes nothing past cap
    assert need == len(b"[0,1,2,3,4,5,6,7]") and buf.raw[:1] == b"["
    big = ctypes.create_string_buffer(need)
    assert L.hb_json_f32_array(v.ctypes.data, v.size, big, need) == need and big.raw == b"[0,1,2,3,4,5,6,7]"


def test_embeddings_route_serialised_natively_equals_the_dict_form(front):
    rt, base = front
    body = {"model": "tiny", "input": [[1, 2, 3], [4, 5, 6, 7, 8]]}
    got = post(base, "/v1/embeddings", body)
    assert got["object"] == "list" and got["model"] == "tiny" and got["usage"] == {"prompt_tokens": 8, "total_tokens": 8}
    assert [d["index"] for d in got["data"]] == [0, 1] and all(d["object"] == "embedding" for d in got["data"])
    assert got["data"][0]["embedding"] == [3.0] * 8 and got["data"][1]["embedding"] == [5.0] * 8


def test_small_responses_are_not_held_up_by_nagle(front):
    """Headers and body sent as two small segments with Nagle on cost one delayed ACK (~40 ms on Linux) per response —
    the embedding route measured 50 ms per request on the GPU box before TCP_NODELAY + a single send."""
    import http.client
    import time
    from urllib.parse import urlparse
    rt, base = front
    u = urlparse(base)
    c = http.client.HTTPConnection(u.hostname, u.port, timeout=30)
    lat = []
    for i in range(30):
        body = json.dumps({"input": [[1, 2, 3, 4 + i]]})
        t0 = time.monotonic()
        c.request("POST", "/v1/embeddings", body, {"Content-Type": "application/json"})
        r = c.getresponse()
        d = json.loads(r.read())
        lat.append(time.monotonic() - t0)
        assert r.status == 200 and d["data"][0]["embedding"] == [4.0] * 8
    lat.sort()
    assert lat[len(lat) // 2] < 0.02, lat   # the batching window is 2 ms; 40 ms means a delayed ACK is in the path


def test_burst_of_concurrent_streams_is_accepted(front):
    """--max-num-seqs 256 means up to 256 streams connect at once: none may be reset by a short listen backlog."""
    import http.client
    import threading
    from urllib.parse import urlparse
    rt, base = front
    u = urlparse(base)
    go, res = threading.Event(), []

    def one():
        go.wait()
        try:
            c = http.client.HTTPConnection(u.hostname, u.port, timeout=60)
            c.request("POST", "/v1/chat/completions", json.dumps({"model": "tiny", "stream": True, "max_tokens": 3,
                                                                    "messages": [{"role": "user", "content": "hi"}]}),
                      {"Content-Type": "application/json"})
            r = c.getresponse()
            body = r.read().decode()
            res.append((r.status, body.rstrip().endswith("data: [DONE]")))
        except OSError as e:
            res.append((repr(e), False))

    ths = [threading.Thread(target=one) for _ in range(256)]
    for t in ths:
        t.start()
    go.set()
    for t in ths:
        t.join()
    assert len(res) == 256 and all(r == (200, True) for r in res), [r for r in res if r != (200, True)][:3]


def test_client_disconnect_mid_stream_cancels_and_releases_the_request(front):
    """The reference's backends stop generating when the HTTP client goes away; here the sequence holds KV pages until
    it is cancelled, so a dropped stream must end in hb_cancel + hb_release, not in a generation nobody reads."""
    import socket
    import time
    from urllib.parse import urlparse
    rt, base = front
    eng = rt.engine
    eng.script = [(b % 26) + 97 + ByteTokenizer.OFFSET for b in range(200000)]   # an answer that never ends on its own
    slow_wait = eng.wait
    eng.wait = lambda rid, ms: (time.sleep(0.005), slow_wait(rid, ms))[1]        # ~200 polls per second
    u = urlparse(base)
    s = socket.create_connection((u.hostname, u.port))
    body = json.dumps({"model": "tiny", "stream": True, "max_tokens": 100000, "messages": [{"role": "user", "content": "go"}]}).encode()
    s.sendall(b"POST /v1/chat/completions HTTP/1.1\r\nHost: x\r\nContent-Type: application/json\r\nContent-Length: %d\r\n\r\n" % len(body) + body)
    got = b""
    while got.count(b"data: ") < 5:
        got += s.recv(65536)
    assert b"200 OK" in got and not eng.cancelled
    s.setsockopt(socket.SOL_SOCKET, socket.SO_LINGER, b"\x01\x00\x00\x00\x00\x00\x00\x00")   # RST on close
    s.close()
    deadline = time.monotonic() + 5
    while time.monotonic() < deadline and not eng.released:
        time.sleep(0.02)
    assert eng.cancelled == [1] and eng.released == [1]
    pos = eng.reqs[1]["pos"]
    time.sleep(0.1)
    assert eng.reqs[1]["pos"] == pos          # nobody polls the cancelled request any more


def test_n_choices_logprobs_and_penalties_reach_the_engine(front):
    """Request fields the runner forwards untouched (openai_chat_handlers.go:100-175): n, logprobs/top_logprobs,
    presence/frequency penalties — one engine submission per choice with its own seed, OpenAI-shaped logprobs back."""
    rt, base = front
    r = post(base, "/v1/chat/completions", {"model": "tiny", "max_tokens": 4, "n": 2, "seed": 5, "logprobs": True,
                                            "top_logprobs": 2, "presence_penalty": 0.5, "frequency_penalty": -0.25,
                                            "messages": [{"role": "user", "content": "hi"}]})
    assert [c["index"] for c in r["choices"]] == [0, 1]
    assert all(c["message"]["content"] == "Hell" and c["finish_reason"] == "length" for c in r["choices"])
    sps = [rt.engine.reqs[i]["sp"] for i in (1, 2)]
    assert [sp.seed for sp in sps] == [5, 6] and all(sp.logprobs == 3 for sp in sps)
    assert all(sp.presence_penalty == 0.5 and sp.frequency_penalty == -0.25 for sp in sps)
    lp = r["choices"][0]["logprobs"]["content"]
    assert [e["token"] for e in lp] == list("Hell") and all(e["logprob"] == -0.5 and len(e["top_logprobs"]) == 2 for e in lp)
    assert lp[0]["bytes"] == [ord("H")] and lp[0]["top_logprobs"][0]["logprob"] == -1.0
    assert r["usage"]["completion_tokens"] == 8 and sorted(rt.engine.released) == [1, 2]
    # streaming: chunks carry the choice index; each choice ends with its own finish_reason, [DONE] closes the stream
    raw = post(base, "/v1/chat/completions", {"model": "tiny", "max_tokens": 4, "n": 2, "stream": True,
                                              "messages": [{"role": "user", "content": "hi"}]}, raw=True)
    chunks = [json.loads(e[6:]) for e in raw.split("\n\n") if e and e != "data: [DONE]"]
    fins = [(c["choices"][0]["index"], c["choices"][0]["finish_reason"]) for c in chunks if c["choices"][0]["finish_reason"]]
    assert sorted(fins) == [(0, "length"), (1, "length")] and "usage" in chunks[-1]
    for bad in ({"n": 0}, {"n": 99}, {"top_logprobs": 21}, {"presence_penalty": 3}):
        with pytest.raises(urllib.error.HTTPError) as e:
            post(base, "/v1/chat/completions", dict({"model": "tiny", "messages": [{"role": "user", "content": "x"}]}, **bad))
        assert e.value.code == 400


def test_completions_prompt_forms_and_id_validation(front):
    rt, base = front
    r = post(base, "/v1/completions", {"model": "tiny", "prompt": "hello", "max_tokens": 3})      # standard string prompt
    assert r["choices"][0]["message"]["content"] == "Hel" and rt.engine.reqs[1]["n_prompt"] == 6  # BOS + 5 bytes
    r = post(base, "/v1/completions", {"model": "tiny", "prompt": [5, 6, 7], "max_tokens": 3})     # token-id array
    assert rt.engine.reqs[2]["n_prompt"] == 3
    for bad in ({"prompt": [5, 100000]}, {"prompt": [1.5]}, {"prompt": None}, {"messages": "hi"}, {}):
        with pytest.raises(urllib.error.HTTPError) as e:
            post(base, "/v1/completions", dict({"model": "tiny"}, **bad))
        assert e.value.code == 400, bad                       # never a dropped connection, never a silent id wrap
    with pytest.raises(urllib.error.HTTPError) as e:
        post(base, "/v1/embeddings", {"model": "tiny", "input": [[5, 100000]]})
    assert e.value.code == 400


def test_stream_errors_are_real_http_errors_and_engine_failures_are_not_stop(front):
    """Validation / submission failures of a stream:true request happen before the 200 + event-stream headers, so the
    client sees a real 4xx/5xx; once the stream is open an engine-side abort becomes an SSE error event without a
    finish_reason chunk or [DONE] — never `finish_reason: "stop"` on a truncated answer."""
    rt, base = front
    with pytest.raises(urllib.error.HTTPError) as e:
        post(base, "/v1/chat/completions", {"model": "other", "stream": True, "messages": []}, raw=True)
    assert e.value.code == 400
    orig = rt.engine.submit

    def full(ids, sp):
        raise HBError(-6, "queue full")
    rt.engine.submit = full
    with pytest.raises(urllib.error.HTTPError) as e:
        post(base, "/v1/chat/completions", {"model": "tiny", "stream": True, "messages": [{"role": "user", "content": "x"}]}, raw=True)
    assert e.value.code == 429
    rt.engine.submit = orig
    polls = {"n": 0}
    orig_poll = rt.engine.poll

    def failing_poll(rid):
        polls["n"] += 1
        if polls["n"] >= 3:
            return [], 2          # engine FAILED / CANCELLED the sequence
        return orig_poll(rid)
    rt.engine.poll = failing_poll
    raw = post(base, "/v1/chat/completions", {"model": "tiny", "stream": True, "max_tokens": 100,
                                              "messages": [{"role": "user", "content": "x"}]}, raw=True)
    events = [json.loads(e[6:]) for e in raw.split("\n\n") if e.startswith("data: {")]
    assert "error" in events[-1] and "data: [DONE]" not in raw
    assert all(ev["choices"][0]["finish_reason"] is None for ev in events[:-1])
    with pytest.raises(urllib.error.HTTPError) as e:   # same failure without streaming: a 500, not a 200 with finish "stop"
        polls["n"] = 0
        post(base, "/v1/chat/completions", {"model": "tiny", "max_tokens": 100, "messages": [{"role": "user", "content": "x"}]})
    assert e.value.code == 500
