"""OpenAI-compatible HTTP front of a B200Runtime (`Runtime.URL()` must serve
POST /v1/chat/completions (JSON and `stream:true` SSE ending in a chunk with a non-empty finish_reason, which is what
closes the control-plane stream — api/pkg/openai/helix_openai_client.go:197), POST /v1/embeddings and GET /v1/models).

Row F1 of the scope table: a small stdlib server, one thread per connection, token ids in/out of the engine C ABI.
Tokenisation is pluggable; without a tokenizer file (no checkpoints offline) a byte-level stand-in is used and
`input` / `prompt` may also be given as token-id arrays (the reference accepts `[][]int`, types/types.go:2707-2730).
"""
import json
import threading
import time
import uuid
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer

import numpy as np

from .engine import HBError, Sampling


class ByteTokenizer:
    """UTF-8 bytes shifted past a few control ids; deterministic stand-in for random-init models."""
    BOS, EOS, OFFSET = 1, 2, 3

    def encode(self, text):
        return [self.BOS] + [b + self.OFFSET for b in text.encode("utf-8")]

    def decode(self, ids):
        return bytes((i - self.OFFSET) % 256 for i in ids if i >= self.OFFSET).decode("utf-8", errors="replace")

    def chat(self, messages):
        return self.encode("".join(f"<|{m.get('role', 'user')}|>\n{m.get('content', '')}\n" for m in messages) + "<|assistant|>\n")


class HFTokenizer:
    """`tokenizers` JSON file + the Llama-3 chat template."""

    def __init__(self, path, eos_token="<|eot_id|>"):
        from tokenizers import Tokenizer
        self.tk = Tokenizer.from_file(path)
        self.EOS = self.tk.token_to_id(eos_token) if self.tk.token_to_id(eos_token) is not None else -1

    def encode(self, text):
        return self.tk.encode(text, add_special_tokens=False).ids

    def decode(self, ids):
        return self.tk.decode(ids)

    def chat(self, messages):
        s = "<|begin_of_text|>"
        for m in messages:
            s += f"<|start_header_id|>{m.get('role', 'user')}<|end_header_id|>\n\n{m.get('content', '')}<|eot_id|>"
        return self.encode(s + "<|start_header_id|>assistant<|end_header_id|>\n\n")


_F32_JSON = []


def native_f32_json():
    """float32 vector -> JSON array bytes through libhelixb200.so (hb_json_f32_array), or None without the library."""
    if not _F32_JSON:
        try:
            import ctypes
            from . import _lib
            L = _lib.lib()

            def fmt(v):
                v = np.ascontiguousarray(v, dtype=np.float32)
                cap = 16 * v.size + 2           # a float32 never needs more than 15 characters + comma
                buf = ctypes.create_string_buffer(cap)
                n = L.hb_json_f32_array(v.ctypes.data, v.size, buf, cap)
                return buf.raw[:n]
            _F32_JSON.append(fmt)
        except Exception:  # noqa: BLE001 — the front also runs against stand-in engines without the shared library
            _F32_JSON.append(None)
    return _F32_JSON[0]


class EmbedBatcher:
    """Server-side coalescing of embedding requests (scope row F4).  The reference's RAG caller sends ONE chunk per
    request with 10 concurrent workers (api/pkg/rag/rag_pgvector.go:70-83) in batches of 50
    (controller/knowledge/knowledge_indexer.go:586-603); encoding them one by one would leave the GPU idle, so requests
    that arrive within `window_s` (or until `max_seqs`) are encoded by a single hb_embed call."""

    def __init__(self, engine, window_s=0.002, max_seqs=256):
        self.engine, self.window_s, self.max_seqs = engine, window_s, max_seqs
        self.cv = threading.Condition()
        self.pending = []   # [seqs, event, result slot]
        self.stop_flag = False
        self.batches = 0
        self.expect = 1     # requests the next batch waits for (at most window_s): the size of the previous one
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def embed(self, seqs):
        item = {"seqs": seqs, "done": threading.Event(), "out": None, "err": None}
        with self.cv:
            self.pending.append(item)
            self.cv.notify()
        item["done"].wait()
        if item["err"] is not None:
            raise item["err"]
        return item["out"]

    def _run(self):
        while True:
            with self.cv:
                while not self.pending and not self.stop_flag:
                    self.cv.wait()
                if self.stop_flag and not self.pending:
                    return
                # wait for company, but not longer than it takes: the callers are a fixed pool of workers that each send
                # their next chunk as soon as the previous answer arrives, so once as many requests are waiting as the
                # last batch held, nobody else is about to show up
                deadline = time.monotonic() + self.window_s
                while sum(len(i["seqs"]) for i in self.pending) < self.max_seqs and len(self.pending) < self.expect:
                    left = deadline - time.monotonic()
                    if left <= 0:
                        break
                    self.cv.wait(left)
                batch, self.pending = self.pending, []
                self.expect = max(1, len(batch))
            flat = [s for i in batch for s in i["seqs"]]
            try:
                vecs = self.engine.embed(flat)
                k = 0
                for i in batch:
                    i["out"] = vecs[k:k + len(i["seqs"])]
                    k += len(i["seqs"])
            except Exception as e:  # one bad sequence must not poison its neighbours: retry each request alone
                for i in batch:
                    try:
                        i["out"] = self.engine.embed(i["seqs"])
                    except Exception as e2:
                        i["err"] = e2
            self.batches += 1
            for i in batch:
                i["done"].set()

    def close(self):
        with self.cv:
            self.stop_flag = True
            self.cv.notify_all()
        self.thread.join(timeout=5)


class StreamDecoder:
    """Token ids -> text for streaming: the visible text must not depend on how the tokens were grouped into polls.
    Decoding each poll's ids on its own garbles any character whose bytes / pieces straddle two polls; decoding the whole
    sequence at every poll is quadratic in the length of the generation.  So (the scheme vLLM's detokenizer uses) the
    tokens not yet shown are decoded together with the previously shown batch as left context, the context's own text is
    cut off the front, and nothing is released while the tail is an incomplete multi-byte sequence (a trailing U+FFFD)
    unless the stream has ended."""

    def __init__(self, tok, skip=()):
        self.tok, self.skip, self.ids = tok, set(skip), []
        self.ctx_off = 0    # first token of the left context
        self.read_off = 0   # tokens before this index have been shown

    def feed(self, ids, final=False):
        self.ids += [t for t in ids if t not in self.skip]
        n = len(self.ids)
        if self.read_off == n:
            return ""
        ctx_text = self.tok.decode(self.ids[self.ctx_off:self.read_off])
        # an unfinished character is at most its last three tokens: release up to the latest clean cut
        for hold in range(0, 1 if final else min(4, n - self.read_off)):
            text = self.tok.decode(self.ids[self.ctx_off:n - hold])
            if final or (not text.endswith("\ufffd") and len(text) > len(ctx_text)):
                self.ctx_off, self.read_off = self.read_off, n - hold
                return text[len(ctx_text):]
        if n - self.read_off > 8:  # not an unfinished character but invalid bytes: nothing later will repair them
            text = self.tok.decode(self.ids[self.ctx_off:])
            self.ctx_off, self.read_off = self.read_off, n
            return text[len(ctx_text):]
        return ""


class StopMatcher:
    """OpenAI `stop` (string or list of up to 4 strings): generation ends at the first occurrence and the stop text is
    not returned.  Streaming: text that could still turn out to be the beginning of a stop string is held back."""

    def __init__(self, stop):
        if stop is None:
            stop = []
        if isinstance(stop, str):
            stop = [stop]
        self.stops = [s for s in stop if s]
        self.hold = max((len(s) for s in self.stops), default=1) - 1
        self.buf = ""
        self.hit = False

    def feed(self, text):
        """Returns the text that may be emitted now."""
        if self.hit:
            return ""
        if not self.stops:
            return text
        self.buf += text
        cut = min((i for i in (self.buf.find(s) for s in self.stops) if i >= 0), default=-1)
        if cut >= 0:
            out, self.buf, self.hit = self.buf[:cut], "", True
            return out
        keep = 0  # longest suffix of buf that is a proper prefix of some stop string
        for k in range(min(self.hold, len(self.buf)), 0, -1):
            if any(s.startswith(self.buf[-k:]) for s in self.stops):
                keep = k
                break
        out, self.buf = self.buf[:len(self.buf) - keep], self.buf[len(self.buf) - keep:]
        return out

    def flush(self):
        out, self.buf = ("" if self.hit else self.buf), ""
        return out


def chat_chunk(cid, model, created, delta, finish_reason, index=0, logprobs=None):
    c = {"index": index, "delta": delta, "finish_reason": finish_reason}
    if logprobs is not None:
        c["logprobs"] = logprobs
    return {"id": cid, "object": "chat.completion.chunk", "created": created, "model": model, "choices": [c]}


class OpenAIServer:
    def __init__(self, runtime, tokenizer=None, host="127.0.0.1", port=0):
        self.rt = runtime
        self.tok = tokenizer or ByteTokenizer()
        self.host, self.port = host, port
        self.httpd = None
        self.batcher = None
        self._active = 0                       # POST handlers in flight (stop() drains them before the engine goes away)
        self._active_cv = threading.Condition()
        self._stopping = False

    def _enter(self):
        with self._active_cv:
            self._active += 1

    def _leave(self):
        with self._active_cv:
            self._active -= 1
            self._active_cv.notify_all()

    # ---- request handlers (pure functions of the parsed body: unit-testable without sockets)
    def models(self):
        return {"object": "list", "data": [{"id": m, "object": "model", "owned_by": "helix-b200"} for m in self.rt.list_models()]}

    def _embed_vectors(self, body):
        inp = body.get("input")
        enc = getattr(self.tok, "encode_for_embedding", self.tok.encode)  # [CLS] ... [SEP] framing where the model has one
        if isinstance(inp, str):
            seqs = [enc(inp)]
        elif isinstance(inp, list) and inp and isinstance(inp[0], int):
            seqs = [inp]
        elif isinstance(inp, list):
            seqs = [enc(x) if isinstance(x, str) else list(x) for x in inp]
        else:
            raise ValueError("input must be a string, a list of strings or token arrays")
        vocab, max_ctx = self.rt.engine.desc.vocab, self.rt.engine.cfg.max_ctx
        for k, s in enumerate(seqs):   # C-level checks: this runs per chunk at the indexer's request rate
            if not isinstance(s, list) or (s and (set(map(type, s)) != {int} or min(s) < 0 or max(s) >= vocab)):
                raise ValueError(f"input holds token ids outside [0, {vocab})")
            if len(s) > max_ctx:
                seqs[k] = s[:max_ctx]
        if self.batcher is None:
            self.batcher = EmbedBatcher(self.rt.engine)
        return self.batcher.embed(seqs), sum(map(len, seqs))

    def embeddings(self, body):
        vecs, n_tok = self._embed_vectors(body)
        return {"object": "list", "model": body.get("model", self.rt.p.model),
                "data": [{"object": "embedding", "index": i, "embedding": v.tolist() if hasattr(v, "tolist") else [float(x) for x in v]}
                         for i, v in enumerate(vecs)],
                "usage": {"prompt_tokens": n_tok, "total_tokens": n_tok}}

    def embeddings_bytes(self, body):
        """The /v1/embeddings response as bytes.  Turning 768 floats into JSON text costs CPython ~0.35 ms per vector —
        more than the GPU needs to compute it — so the vectors are written by the library (hb_json_f32_array: shortest
        round-trip decimal per float32, what Go's encoding/json emits) and only the envelope is built here."""
        vecs, n_tok = self._embed_vectors(body)
        fmt = native_f32_json()
        if fmt is None or not hasattr(vecs, "dtype"):
            return json.dumps(self.embeddings_from(vecs, n_tok, body)).encode()
        parts = [b'{"object":"list","model":', json.dumps(body.get("model", self.rt.p.model)).encode(), b',"data":[']
        for i, v in enumerate(vecs):
            parts.append((b"," if i else b"") + b'{"object":"embedding","index":%d,"embedding":' % i)
            parts.append(fmt(v))
            parts.append(b"}")
        parts.append(b'],"usage":{"prompt_tokens":%d,"total_tokens":%d}}' % (n_tok, n_tok))
        return b"".join(parts)

    def embeddings_from(self, vecs, n_tok, body):
        return {"object": "list", "model": body.get("model", self.rt.p.model),
                "data": [{"object": "embedding", "index": i, "embedding": v.tolist() if hasattr(v, "tolist") else [float(x) for x in v]}
                         for i, v in enumerate(vecs)],
                "usage": {"prompt_tokens": n_tok, "total_tokens": n_tok}}

    def _parse_chat(self, body):
        """Validates the request and returns (prompt ids, [Sampling per choice]).  Raises ValueError -> HTTP 400."""
        if body.get("model") and body["model"] != self.rt.p.model:
            raise ValueError(f"model mismatch, expecting {self.rt.p.model}")  # openai_chat_handlers.go:44-50
        msgs = body.get("messages")
        if msgs is not None:
            if not isinstance(msgs, list) or not all(isinstance(m, dict) for m in msgs):
                raise ValueError("messages must be a list of {role, content} objects")
            ids = self.tok.chat(msgs)
        else:
            prompt = body.get("prompt")
            if isinstance(prompt, str):          # /v1/completions with the standard string prompt
                ids = self.tok.encode(prompt)
            elif isinstance(prompt, list) and prompt and all(isinstance(t, int) and not isinstance(t, bool) for t in prompt):
                ids = list(prompt)               # token-id array form
            else:
                raise ValueError("request needs `messages`, or `prompt` as a string or an array of token ids")
        vocab = self.rt.engine.desc.vocab
        if not ids or any(t < 0 or t >= vocab for t in ids):
            raise ValueError(f"prompt is empty or holds token ids outside [0, {vocab})")
        temp = body.get("temperature", 0.0) or 0.0   # the runner already rewrote 0 -> 0.1 (openai_chat_handlers.go:52-58)
        n = body.get("n")
        n = 1 if n is None else n
        if not isinstance(n, int) or isinstance(n, bool) or n < 1 or n > 16:
            raise ValueError("n must be an integer in [1, 16]")
        top_lp = body.get("top_logprobs")
        want_lp = bool(body.get("logprobs")) or top_lp is not None
        if top_lp is not None and not (isinstance(top_lp, int) and 0 <= top_lp <= 20):
            raise ValueError("top_logprobs must be an integer in [0, 20]")
        pres, freq = float(body.get("presence_penalty") or 0.0), float(body.get("frequency_penalty") or 0.0)
        if not (-2.0 <= pres <= 2.0 and -2.0 <= freq <= 2.0):
            raise ValueError("presence_penalty / frequency_penalty must be in [-2, 2]")
        seed = int(body.get("seed", 0) or 0)
        sps = [Sampling(temperature=float(temp), seed=seed + i,   # n > 1: one noise stream per choice
                        max_tokens=int(body.get("max_tokens") or body.get("max_completion_tokens") or 256),
                        eos_token=getattr(self.tok, "EOS", -1),
                        top_p=float(body.get("top_p") or 1.0),      # openai.ChatCompletionRequest.TopP (nucleus)
                        top_k=int(body.get("top_k") or 0),          # vLLM extension the reference's backend accepts
                        logprobs=(1 + int(top_lp or 0)) if want_lp else 0,
                        presence_penalty=pres, frequency_penalty=freq) for i in range(n)]
        return ids, sps

    def _lp_entry(self, ids, lps):
        """One OpenAI `logprobs.content[]` element from a row of hb_logprobs (column 0 = the sampled token)."""
        def one(t, lp):
            text = self.tok.decode([int(t)]) if t >= 0 else ""
            return {"token": text, "logprob": float(max(lp, -9999.0)), "bytes": list(text.encode("utf-8"))}
        e = one(ids[0], lps[0])
        e["top_logprobs"] = [one(t, lp) for t, lp in zip(ids[1:], lps[1:]) if t >= 0]
        return e

    def chat_stream(self, body):
        """Generator of SSE chunk dicts; the last chunk of every choice carries its finish_reason, the very last one the
        usage.  Validation and submission happen BEFORE the first yield (`next()` on the generator raises), so the HTTP
        layer can still answer 4xx/5xx instead of an already-open 200 stream."""
        eng = self.rt.engine
        ids, sps = self._parse_chat(body)
        rids = []
        try:
            for sp in sps:
                rids.append(eng.submit(ids, sp))
        except Exception:
            for r in rids:
                self._retire(r, finished=False)
            raise
        return self._stream_choices(body, ids, sps, rids)

    def _retire(self, rid, finished):
        eng = self.rt.engine
        try:
            if not finished:
                eng.cancel(rid)      # stop string hit / client gone / error: free the sequence's KV pages
                for _ in range(200):  # the step loop retires it at the next step boundary
                    if eng.poll(rid)[1]:
                        break
                    eng.wait(rid, 10)
            eng.release(rid)
        except HBError:
            pass

    def _stream_choices(self, body, ids, sps, rids):
        eng = self.rt.engine
        n_prompt = len(ids)
        cid, created, model = "chatcmpl-" + uuid.uuid4().hex[:24], int(time.time()), self.rt.p.model
        ch = [{"rid": r, "sp": sp, "n": 0, "fin": 0, "done": False, "lp_row": 0,
               "stop": StopMatcher(body.get("stop")), "dec": StreamDecoder(self.tok, skip=[sp.eos_token])}
              for r, sp in zip(rids, sps)]
        try:
            for i in range(len(ch)):
                yield chat_chunk(cid, model, created, {"role": "assistant", "content": ""}, None, i)
            while not all(c["done"] for c in ch):
                first = next(c for c in ch if not c["done"])
                eng.wait(first["rid"], 30000 if len(ch) == 1 else 5)
                for i, c in enumerate(ch):
                    if c["done"]:
                        continue
                    toks, fin = eng.poll(c["rid"])
                    if not toks and not fin:
                        continue
                    c["n"] += len(toks)
                    c["fin"] = fin
                    text = c["stop"].feed(c["dec"].feed(toks, final=bool(fin)))
                    lp = None
                    if c["sp"].logprobs and toks:
                        lid, lpv = eng.logprobs(c["rid"], c["lp_row"], len(toks))
                        c["lp_row"] += len(lid)
                        lp = {"content": [self._lp_entry(a, b) for a, b in zip(lid, lpv)]}
                    if text or lp:
                        yield chat_chunk(cid, model, created, {"content": text}, None, i, lp)
                    if fin or c["stop"].hit:
                        tail = c["stop"].flush()
                        if tail:
                            yield chat_chunk(cid, model, created, {"content": tail}, None, i)
                        if fin == 2 and not c["stop"].hit:
                            # the engine FAILED or CANCELLED the sequence: not a normal completion
                            raise HBError(-2, "generation aborted by the engine: " + str(self.rt.status() or "engine stopped"))
                        reason = "stop" if (c["stop"].hit or c["n"] < c["sp"].max_tokens) else "length"
                        c["done"] = True
                        self._retire(c["rid"], finished=bool(fin))   # a stop hit frees the sequence's KV pages now
                        c["retired"] = True
                        last = chat_chunk(cid, model, created, {}, reason, i)
                        if all(x["done"] for x in ch):
                            done_toks = sum(x["n"] for x in ch)
                            last["usage"] = {"prompt_tokens": n_prompt, "completion_tokens": done_toks,
                                             "total_tokens": n_prompt + done_toks}
                        yield last
        finally:
            for c in ch:
                if not c.get("retired"):
                    self._retire(c["rid"], finished=bool(c["fin"]))

    def chat(self, body):
        texts, reasons, lps = {}, {}, {}
        cid, usage = None, None
        for chk in self.chat_stream(body):
            cid = chk["id"]
            c = chk["choices"][0]
            i = c["index"]
            texts[i] = texts.get(i, "") + (c["delta"].get("content", "") or "")
            if c.get("logprobs"):
                lps.setdefault(i, []).extend(c["logprobs"]["content"])
            if c["finish_reason"]:
                reasons[i] = c["finish_reason"]
            usage = chk.get("usage", usage)
        choices = []
        for i in sorted(texts):
            one = {"index": i, "message": {"role": "assistant", "content": texts[i]}, "finish_reason": reasons.get(i, "stop")}
            if i in lps:
                one["logprobs"] = {"content": lps[i]}
            choices.append(one)
        return {"id": cid, "object": "chat.completion", "created": int(time.time()), "model": self.rt.p.model,
                "choices": choices, "usage": usage or {"prompt_tokens": 0, "completion_tokens": 0, "total_tokens": 0}}

    # ---- socket plumbing
    def start(self):
        srv = self

        class H(BaseHTTPRequestHandler):
            protocol_version = "HTTP/1.1"
            # headers and body (or one SSE event after another) are separate small sends: with Nagle on, the second one
            # waits for the client's delayed ACK of the first — 40 ms per response / per token on Linux
            disable_nagle_algorithm = True

            def log_message(self, *a):
                pass

            def _json(self, code, obj, raw=None):
                data = raw if raw is not None else json.dumps(obj).encode()
                self.send_response(code)
                self.send_header("Content-Type", "application/json")
                self.send_header("Content-Length", str(len(data)))
                # one send for headers + body (end_headers() would flush the header block on its own)
                pending = getattr(self, "_headers_buffer", None)
                if isinstance(pending, list) and pending:
                    self._headers_buffer = []
                    self.wfile.write(b"".join(pending) + b"\r\n" + data)
                else:   # a stdlib whose handler buffers differently: two sends, still correct (TCP_NODELAY is set)
                    self.end_headers()
                    self.wfile.write(data)

            def do_GET(self):
                if self.path.rstrip("/") in ("/v1/models", "/models"):
                    self._json(200, srv.models())
                elif self.path == "/healthz":
                    self._json(200 if srv.rt.status() else 503, {"status": srv.rt.status()})
                else:
                    self._json(404, {"error": "not found"})

            def do_POST(self):
                srv._enter()
                headers_sent = False
                try:
                    n = int(self.headers.get("Content-Length", "0"))
                    if n > 10 * 1024 * 1024:   # openai_chat_handlers.go:40
                        return self._json(413, {"error": "request too large"})
                    body = json.loads(self.rfile.read(n) or b"{}")
                    if not isinstance(body, dict):
                        raise ValueError("request body must be a JSON object")
                    path = self.path.rstrip("/")
                    if path.endswith("/embeddings"):
                        return self._json(200, None, raw=srv.embeddings_bytes(body))
                    if path.endswith("/chat/completions") or path.endswith("/completions"):
                        if not body.get("stream"):
                            return self._json(200, srv.chat(body))
                        # validation + submission run before any byte of the response: errors are still real 4xx/5xx
                        chunks = srv.chat_stream(body)
                        first = next(chunks)
                        self.send_response(200)
                        self.send_header("Content-Type", "text/event-stream")
                        self.send_header("Cache-Control", "no-cache")
                        self.send_header("Connection", "close")
                        self.end_headers()
                        headers_sent = True
                        self.close_connection = True
                        self.wfile.write(b"data: " + json.dumps(first).encode() + b"\n\n")
                        for ch in chunks:
                            self.wfile.write(b"data: " + json.dumps(ch).encode() + b"\n\n")
                            self.wfile.flush()
                        self.wfile.write(b"data: [DONE]\n\n")
                        self.wfile.flush()
                        return
                    self._json(404, {"error": "not found"})
                except (BrokenPipeError, ConnectionResetError):
                    pass
                except Exception as e:  # noqa: BLE001 — every failure must reach the client as an error, never a dropped socket
                    bad_request = isinstance(e, (ValueError, KeyError, TypeError)) or (isinstance(e, HBError) and e.code in (-1, -5))
                    err = {"error": {"message": str(e), "type": "invalid_request_error" if bad_request else "server_error"}}
                    try:
                        if headers_sent:   # the 200 is out: an SSE error event, then the connection closes without [DONE]
                            self.wfile.write(b"data: " + json.dumps(err).encode() + b"\n\n")
                            self.wfile.flush()
                        else:
                            busy = isinstance(e, HBError) and e.code == -6
                            self._json(400 if bad_request else (429 if busy else 500), err)
                    except (BrokenPipeError, ConnectionResetError):
                        pass
                finally:
                    srv._leave()

        class Server(ThreadingHTTPServer):
            # the slot's default concurrency is 256 streams (--max-num-seqs): with socketserver's backlog of 5 a burst of
            # connections is reset by the kernel before accept() gets to them
            request_queue_size = 1024
            daemon_threads = True

        self.httpd = Server((self.host, self.port), H)
        self.httpd.daemon_threads = True
        self.port = self.httpd.server_address[1]
        threading.Thread(target=self.httpd.serve_forever, daemon=True).start()
        return f"http://{self.host}:{self.port}"

    def stop(self, drain_s=10.0):
        """Stops accepting connections and waits for in-flight handlers: they hold engine handles, and Runtime.Stop
        destroys the engine right after this returns.  hb_engine_destroy wakes anything still parked in hb_wait."""
        self._stopping = True
        if self.httpd:
            self.httpd.shutdown()
        deadline = time.monotonic() + drain_s
        with self._active_cv:
            while self._active > 0 and time.monotonic() < deadline:
                self._active_cv.wait(0.05)
        if self.batcher:
            self.batcher.close()
            self.batcher = None
        if self.httpd:
            self.httpd.shutdown()
            self.httpd.server_close()
            self.httpd = None
