"""Executable mirror of integration/patches/0001-runner-nats-relay-coalesce-sse-lines.patch (scope row F1, relay half).

The reference's runner relays a streaming backend response over NATS one SSE line per publish
(api/pkg/runner/controller_nats.go:235-274) and the control plane parses exactly one chunk per message
(api/pkg/openai/helix_openai_client.go:156-205).  The patch coalesces every COMPLETE line that is already buffered into
one publish and makes the consumer accept several newline-separated chunks; this module restates both halves so the
algorithm is tested (tests/test_relay_cpu.py) — the Go patch itself cannot be compiled in this image."""
import json

MAX_BATCH_BYTES = 256 * 1024   # well under NATS_SERVER_MAX_PAYLOAD (api/pkg/config/config.go:424-429)


class BufferedLines:
    """What the patch needs from bufio.Reader: a blocking read of one line, and a non-blocking look at buffered bytes."""

    def __init__(self, read_some):
        self._read_some, self._buf, self._eof = read_some, b"", False

    def _fill(self):
        data = self._read_some()          # blocks until the backend writes something (b"" = end of stream)
        if not data:
            self._eof = True
        self._buf += data

    def read_line(self):
        """bufio.Reader.ReadBytes('\\n'): blocks until a full line (or EOF) is available."""
        while b"\n" not in self._buf and not self._eof:
            self._fill()
        if b"\n" in self._buf:
            line, self._buf = self._buf.split(b"\n", 1)
            return line + b"\n"
        line, self._buf = self._buf, b""
        return line                       # may be b"" at EOF

    def buffered_line(self):
        """A complete line that is ALREADY buffered, or None — never waits for the backend (Peek + IndexByte in the patch)."""
        if b"\n" not in self._buf:
            return None
        line, self._buf = self._buf.split(b"\n", 1)
        return line + b"\n"


def relay_stream(reader: BufferedLines, publish):
    """handleStreamingResponse with coalescing: one publish per batch of complete, already buffered SSE lines."""
    publishes = 0
    while True:
        chunk = reader.read_line()
        if not chunk:
            return publishes
        if not chunk.strip():
            continue                                   # SSE event separator
        batch = chunk.rstrip(b"\r\n")
        while len(batch) < MAX_BATCH_BYTES:
            line = reader.buffered_line()
            if line is None:
                break                                  # an incomplete line stays for the next blocking read
            line = line.strip()
            if not line:
                continue
            batch += b"\n" + line
        publish(batch)
        publishes += 1


def consume_message(payload: bytes):
    """The patched consumer: every `data: {...}` line of a message becomes one chunk; returns (chunks, done)."""
    chunks, done = [], False
    for line in payload.decode().split("\n"):
        body = line.strip()
        if body.startswith("data: "):
            body = body[len("data: "):]
        if not body or body == "[DONE]":
            continue
        obj = json.loads(body)
        chunks.append(obj)
        if obj.get("choices") and obj["choices"][0].get("finish_reason"):
            done = True
            break
    return chunks, done
