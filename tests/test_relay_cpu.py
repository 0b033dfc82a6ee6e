"""CPU: the NATS relay coalescing of integration/patches/0001-* (helix_b200/relay.py is its executable mirror): whatever
the arrival pattern of the backend's bytes, the control plane sees the same chunks in the same order, a burst costs one
publish instead of one per SSE line, nothing waits for data that has not arrived, and old one-line messages still parse."""
import json

from hypothesis import given, settings, strategies as st

from helix_b200.relay import BufferedLines, consume_message, relay_stream
from helix_b200.server import chat_chunk


def sse_bytes(n_tokens):
    out = b""
    for i in range(n_tokens):
        out += b"data: " + json.dumps(chat_chunk("id", "m", 0, {"content": f"t{i}"}, None)).encode() + b"\n\n"
    out += b"data: " + json.dumps(chat_chunk("id", "m", 0, {}, "stop")).encode() + b"\n\n"
    return out + b"data: [DONE]\n\n"


def run(stream, cuts):
    pieces, i, k = [], 0, 0
    while i < len(stream):
        n = cuts[k % len(cuts)]
        k += 1
        pieces.append(stream[i:i + n])
        i += n
    it = iter(pieces)
    published = []
    n_pub = relay_stream(BufferedLines(lambda: next(it, b"")), published.append)
    chunks, done = [], False
    for msg in published:
        c, d = consume_message(msg)
        chunks += c
        done = done or d
    return n_pub, chunks, done, published


def test_burst_is_one_publish_and_trickle_is_one_per_line():
    stream = sse_bytes(50)
    n_pub, chunks, done, _ = run(stream, [len(stream)])            # everything already buffered: one publish
    assert n_pub == 1 and done and [c["choices"][0]["delta"].get("content") for c in chunks[:-1]] == [f"t{i}" for i in range(50)]
    n_pub, chunks2, done, msgs = run(stream, [1])                  # byte by byte: nothing to coalesce, never blocks on a partial line
    assert n_pub == 52 and done and chunks2 == chunks              # 50 tokens + finish + [DONE]
    assert all(b"\n" not in m for m in msgs)
    old = b"data: " + json.dumps(chat_chunk("id", "m", 0, {"content": "x"}, None)).encode()
    assert consume_message(old) == ([json.loads(old[6:])], False)  # an unpatched runner's message is still valid input


@settings(max_examples=300, deadline=None)
@given(st.integers(0, 40), st.lists(st.integers(1, 400), min_size=1, max_size=30))
def test_any_arrival_pattern_yields_the_same_chunks(n_tokens, cuts):
    stream = sse_bytes(n_tokens)
    ref = run(stream, [len(stream)])[1]
    n_pub, chunks, done, _ = run(stream, cuts)
    assert chunks == ref and done and 1 <= n_pub <= n_tokens + 2
