"""ctypes binding of the native tokenizer (hb_tok_*, csrc/tokenizer.cpp) with the interface helix_b200/server.py expects
from a tokenizer: encode / decode / chat / EOS.  The Go shim binds the same symbols through cgo."""
import ctypes as C

import numpy as np

from . import _lib
from .engine import HBError


class NativeTokenizer:
    def __init__(self, tokenizer_json_path, eos_token="<|eot_id|>"):
        self._l = _lib.lib()
        self._h = C.c_void_p()
        rc = self._l.hb_tok_load(str(tokenizer_json_path).encode(), C.byref(self._h))
        if rc != 0:
            raise HBError(rc, f"hb_tok_load({tokenizer_json_path}) failed")
        self.EOS = self._l.hb_tok_token_id(self._h, eos_token.encode())
        self.vocab_size = self._l.hb_tok_vocab_size(self._h)

    def close(self):
        if self._h:
            self._l.hb_tok_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def token_id(self, token):
        return self._l.hb_tok_token_id(self._h, token.encode())

    def encode_for_embedding(self, text):
        """The model's own framing around the text ([CLS] ... [SEP] for WordPiece encoders): what the backend's tokenizer
        adds before an embedding forward pass."""
        return self.encode(text, parse_special=2)

    def encode(self, text, parse_special=False):
        raw = text.encode("utf-8")
        cap = max(16, len(raw) * 2 + 8)  # WordPiece: Chinese characters are isolated, accents may decompose
        n = C.c_int32()
        buf = np.empty(cap, np.int32)
        rc = self._l.hb_tok_encode(self._h, raw, int(parse_special), buf.ctypes.data, cap, C.byref(n))
        if rc == -6:  # HB_ERR_BUSY: buffer too small (cannot happen: a token covers >= 1 byte)
            buf = np.empty(n.value, np.int32)
            rc = self._l.hb_tok_encode(self._h, raw, int(parse_special), buf.ctypes.data, n.value, C.byref(n))
        if rc != 0:
            raise HBError(rc, "hb_tok_encode failed")
        return buf[:n.value].tolist()

    def decode_bytes(self, ids, skip_special=True):
        a = np.ascontiguousarray(ids, dtype=np.int32)
        ln = C.c_size_t()
        cap = 16 + 64 * max(1, a.size)
        out = C.create_string_buffer(cap)
        rc = self._l.hb_tok_decode(self._h, a.ctypes.data, a.size, int(skip_special), out, cap, C.byref(ln))
        if rc == -6:
            out = C.create_string_buffer(ln.value + 1)
            rc = self._l.hb_tok_decode(self._h, a.ctypes.data, a.size, int(skip_special), out, ln.value + 1, C.byref(ln))
        if rc != 0:
            raise HBError(rc, "hb_tok_decode failed")
        return out.raw[:ln.value]

    def decode(self, ids, skip_special=True):
        return self.decode_bytes(ids, skip_special).decode("utf-8", errors="replace")

    def chat(self, messages):
        n_msgs = len(messages)
        roles = (C.c_char_p * max(1, n_msgs))(*[str(m.get("role", "user")).encode() for m in messages])
        conts = (C.c_char_p * max(1, n_msgs))(*[str(m.get("content", "") or "").encode() for m in messages])
        n = C.c_int32()
        cap = 64 + sum(len(m.get("content", "") or "".encode()) + 16 for m in messages) * 2
        buf = np.empty(cap, np.int32)
        rc = self._l.hb_tok_chat_llama3(self._h, roles, conts, n_msgs, buf.ctypes.data, cap, C.byref(n))
        if rc == -6:
            buf = np.empty(n.value, np.int32)
            rc = self._l.hb_tok_chat_llama3(self._h, roles, conts, n_msgs, buf.ctypes.data, n.value, C.byref(n))
        if rc != 0:
            raise HBError(rc, "hb_tok_chat_llama3 failed (not a Llama-3 style vocabulary?)")
        return buf[:n.value].tolist()
