"""Test infrastructure: a numpy restatement of ggml's block formats (dequantize_row_* in ggml-quants.c, ggml v0.9 / llama.cpp
b6xxx as bundled by Ollama v0.13 — the reference's backend, Dockerfile.runner:111) and a minimal GGUF v3 writer.
Quantised test tensors are RANDOM BLOCK BYTES (every byte pattern is a valid block); their numpy dequantisation is the
ground truth the C++ reader (helix_b200/csrc/gguf.cpp) must reproduce bit for bit."""
import struct

import numpy as np

GGML = {"F32": 0, "F16": 1, "Q4_0": 2, "Q4_1": 3, "Q5_0": 6, "Q5_1": 7, "Q8_0": 8, "Q4_K": 12, "Q5_K": 13, "Q6_K": 14, "BF16": 30}
BLOCK = {"F32": (1, 4), "F16": (1, 2), "BF16": (1, 2), "Q4_0": (32, 18), "Q4_1": (32, 20), "Q5_0": (32, 22), "Q5_1": (32, 24),
         "Q8_0": (32, 34), "Q4_K": (256, 144), "Q5_K": (256, 176), "Q6_K": (256, 210)}


def _f16(b):
    return np.frombuffer(b, dtype="<f2").astype(np.float32)


def _scale_min_k4(j, q):
    if j < 4:
        return int(q[j] & 63), int(q[j + 4] & 63)
    return int((q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4)), int((q[j + 4] >> 4) | ((q[j] >> 6) << 4))


def dequant(kind, raw, n):
    """raw bytes of n elements in format `kind` -> float32[n]."""
    raw = np.frombuffer(raw, dtype=np.uint8)
    if kind == "F32":
        return raw.view("<f4").astype(np.float32).copy()
    if kind == "F16":
        return raw.view("<f2").astype(np.float32)
    if kind == "BF16":
        return (raw.view("<u2").astype(np.uint32) << 16).view(np.float32)
    bs, bb = BLOCK[kind]
    blocks = raw.reshape(n // bs, bb)
    out = np.empty((n // bs, bs), np.float32)
    for i, b in enumerate(blocks):
        if kind == "Q4_0":
            d = _f16(b[:2].tobytes())[0]
            qs = b[2:18].astype(np.int32)
            out[i, :16] = ((qs & 0xF) - 8).astype(np.float32) * d
            out[i, 16:] = ((qs >> 4) - 8).astype(np.float32) * d
        elif kind == "Q4_1":
            d, m = _f16(b[:4].tobytes())
            qs = b[4:20].astype(np.int32)
            out[i, :16] = (qs & 0xF).astype(np.float32) * d + m
            out[i, 16:] = (qs >> 4).astype(np.float32) * d + m
        elif kind in ("Q5_0", "Q5_1"):
            off = 2 if kind == "Q5_0" else 4
            d = _f16(b[:2].tobytes())[0]
            m = _f16(b[2:4].tobytes())[0] if kind == "Q5_1" else np.float32(0)
            qh = int(np.frombuffer(b[off:off + 4].tobytes(), "<u4")[0])
            qs = b[off + 4:off + 20].astype(np.int32)
            j = np.arange(16)
            h0 = ((qh >> j) << 4) & 0x10
            h1 = (qh >> (j + 12)) & 0x10
            x0, x1 = (qs & 0xF) | h0, (qs >> 4) | h1
            if kind == "Q5_0":
                out[i, :16] = (x0 - 16).astype(np.float32) * d
                out[i, 16:] = (x1 - 16).astype(np.float32) * d
            else:
                out[i, :16] = x0.astype(np.float32) * d + m
                out[i, 16:] = x1.astype(np.float32) * d + m
        elif kind == "Q8_0":
            d = _f16(b[:2].tobytes())[0]
            out[i] = b[2:34].view(np.int8).astype(np.float32) * d
        elif kind in ("Q4_K", "Q5_K"):
            d, dmin = _f16(b[:4].tobytes())
            sc = b[4:16]
            if kind == "Q4_K":
                q, qh = b[16:144].astype(np.int32), None
            else:
                qh, q = b[16:48].astype(np.int32), b[48:176].astype(np.int32)
            y, is_, u1, u2, qo = out[i], 0, 1, 2, 0
            for j in range(0, 256, 64):
                s1, m1 = _scale_min_k4(is_, sc)
                s2, m2 = _scale_min_k4(is_ + 1, sc)
                d1, mm1 = np.float32(d * np.float32(s1)), np.float32(dmin * np.float32(m1))
                d2, mm2 = np.float32(d * np.float32(s2)), np.float32(dmin * np.float32(m2))
                lo, hi = q[qo:qo + 32] & 0xF, q[qo:qo + 32] >> 4
                if qh is not None:
                    lo = lo + np.where(qh & u1, 16, 0)
                    hi = hi + np.where(qh & u2, 16, 0)
                y[j:j + 32] = d1 * lo.astype(np.float32) - mm1
                y[j + 32:j + 64] = d2 * hi.astype(np.float32) - mm2
                qo += 32
                is_ += 2
                u1 <<= 2
                u2 <<= 2
        elif kind == "Q6_K":
            ql, qh = b[:128].astype(np.int32), b[128:192].astype(np.int32)
            sc = b[192:208].view(np.int8).astype(np.int32)
            d = _f16(b[208:210].tobytes())[0]
            y = out[i]
            for n0, lo, ho, so in ((0, 0, 0, 0), (128, 64, 32, 8)):
                l = np.arange(32)
                is_ = l // 16
                q1 = ((ql[lo + l] & 0xF) | (((qh[ho + l] >> 0) & 3) << 4)) - 32
                q2 = ((ql[lo + l + 32] & 0xF) | (((qh[ho + l] >> 2) & 3) << 4)) - 32
                q3 = ((ql[lo + l] >> 4) | (((qh[ho + l] >> 4) & 3) << 4)) - 32
                q4 = ((ql[lo + l + 32] >> 4) | (((qh[ho + l] >> 6) & 3) << 4)) - 32
                for k, qq in enumerate((q1, q2, q3, q4)):
                    scale = (d * sc[so + is_ + 2 * k].astype(np.float32)).astype(np.float32)   # ggml: d * sc * q, left to right
                    y[n0 + 32 * k + l] = scale * qq.astype(np.float32)
    return out.reshape(-1)


def random_blocks(kind, n, rng, scale=0.02):
    """Random valid bytes for n elements: fp16 scale fields are drawn small and finite, everything else uniformly."""
    bs, bb = BLOCK[kind]
    if kind in ("F32", "F16", "BF16"):
        x = (rng.standard_normal(n) * scale).astype(np.float32)
        if kind == "F32":
            return x.astype("<f4").tobytes()
        if kind == "F16":
            return x.astype("<f2").tobytes()
        u = x.view(np.uint32)
        return (((u + (((u >> 16) & 1) + 0x7FFF)) >> 16).astype("<u2")).tobytes()
    raw = rng.integers(0, 256, size=(n // bs, bb), dtype=np.uint8)
    def put_f16(col, vals):
        raw[:, col:col + 2] = np.frombuffer(vals.astype("<f2").tobytes(), np.uint8).reshape(-1, 2)
    amp = {"Q8_0": scale / 64, "Q4_0": scale / 4, "Q4_1": scale / 8, "Q5_0": scale / 8, "Q5_1": scale / 16, "Q4_K": scale / 200,
           "Q5_K": scale / 400, "Q6_K": scale / 1000}[kind]
    d = (rng.uniform(0.5, 1.5, n // bs) * amp).astype(np.float32)
    if kind in ("Q4_0", "Q5_0", "Q8_0"):
        put_f16(0, d)
    elif kind in ("Q4_1", "Q5_1"):
        put_f16(0, d)
        put_f16(2, -d * (8 if kind == "Q4_1" else 16))
    elif kind in ("Q4_K", "Q5_K"):
        put_f16(0, d)
        put_f16(2, d * 8)
    elif kind == "Q6_K":
        put_f16(208, d)
    return raw.tobytes()


def _s(x):
    b = x.encode()
    return struct.pack("<Q", len(b)) + b


def write_gguf(path, meta, tensors, align=32):
    """meta: {key: int | float | str}; tensors: [(name, kind, (rows, cols) or (n,), raw bytes)]."""
    out = [b"GGUF", struct.pack("<IQQ", 3, len(tensors), len(meta))]
    for k, v in meta.items():
        out.append(_s(k))
        if isinstance(v, str):
            out.append(struct.pack("<I", 8) + _s(v))
        elif isinstance(v, float):
            out.append(struct.pack("<If", 6, v))
        else:
            out.append(struct.pack("<II", 4, int(v)))
    off, blobs = 0, []
    for name, kind, shape, raw in tensors:
        dims = list(reversed(shape))  # ne[0] = contiguous dimension
        out.append(_s(name) + struct.pack("<I", len(dims)) + b"".join(struct.pack("<Q", d) for d in dims) +
                   struct.pack("<IQ", GGML[kind], off))
        pad = (-len(raw)) % align
        blobs.append(raw + b"\0" * pad)
        off += len(raw) + pad
    head = b"".join(out)
    head += b"\0" * ((-len(head)) % align)
    with open(path, "wb") as f:
        f.write(head)
        for b in blobs:
            f.write(b)


def hf_permute(w, n_head):
    """llama.cpp convert_hf_to_gguf.py LlamaModel.permute: HF q/k rows -> the interleaved-pair order GGUF stores."""
    return w.reshape(n_head, 2, w.shape[0] // n_head // 2, *w.shape[1:]).swapaxes(1, 2).reshape(w.shape)


def llama_gguf(path, d, rng, kinds, arch="llama"):
    """A whole random model: returns the HF-named fp32 state dict the file DEQUANTISES to (bf16-rounded by the loader)."""
    H, F, V, D = d.hidden, d.ffn, d.vocab, d.head_dim
    meta = {"general.architecture": arch, "general.alignment": 32, f"{arch}.block_count": d.layers, f"{arch}.embedding_length": H,
            f"{arch}.feed_forward_length": F, f"{arch}.attention.head_count": d.heads, f"{arch}.attention.head_count_kv": d.kv_heads,
            f"{arch}.attention.layer_norm_rms_epsilon": float(d.norm_eps), f"{arch}.rope.freq_base": float(d.rope_theta),
            f"{arch}.context_length": d.max_pos, f"{arch}.attention.key_length": D}
    tensors, sd = [], {}

    def add(gname, hfname, shape, kind, head_perm=0, ones=False):
        n = int(np.prod(shape))
        bs = BLOCK[kind][0]
        if shape[-1] % bs:
            kind = "F32"
        raw = np.ones(n, "<f4").tobytes() if ones else random_blocks(kind, n, rng)
        w = dequant(kind, raw, n).reshape(shape)
        tensors.append((gname, kind, shape, raw))
        # the file holds the PERMUTED rows for llama-arch q/k; the HF-layout truth is the inverse permutation
        if head_perm and arch == "llama":
            nh = head_perm
            w = w.reshape(nh, shape[0] // nh // 2, 2, *shape[1:]).swapaxes(1, 2).reshape(shape)
        sd[hfname] = w
    add("token_embd.weight", "model.embed_tokens.weight", (V, H), kinds.get("embd", "Q8_0"))
    for i in range(d.layers):
        g, p = f"blk.{i}.", f"model.layers.{i}."
        add(g + "attn_norm.weight", p + "input_layernorm.weight", (H,), "F32", ones=True)
        add(g + "attn_q.weight", p + "self_attn.q_proj.weight", (d.heads * D, H), kinds.get("attn", "Q4_0"), head_perm=d.heads)
        add(g + "attn_k.weight", p + "self_attn.k_proj.weight", (d.kv_heads * D, H), kinds.get("attn", "Q4_0"), head_perm=d.kv_heads)
        add(g + "attn_v.weight", p + "self_attn.v_proj.weight", (d.kv_heads * D, H), kinds.get("v", "Q6_K"))
        if getattr(d, "qkv_bias", 0):
            for t, nrow in (("q", d.heads * D), ("k", d.kv_heads * D), ("v", d.kv_heads * D)):
                add(g + f"attn_{t}.bias", p + f"self_attn.{t}_proj.bias", (nrow,), "F32")
        add(g + "attn_output.weight", p + "self_attn.o_proj.weight", (H, d.heads * D), kinds.get("attn", "Q4_0"))
        add(g + "ffn_norm.weight", p + "post_attention_layernorm.weight", (H,), "F32", ones=True)
        add(g + "ffn_gate.weight", p + "mlp.gate_proj.weight", (F, H), kinds.get("ffn", "Q4_K"))
        add(g + "ffn_up.weight", p + "mlp.up_proj.weight", (F, H), kinds.get("ffn", "Q4_K"))
        add(g + "ffn_down.weight", p + "mlp.down_proj.weight", (H, F), kinds.get("down", "Q5_K"))
    add("output_norm.weight", "model.norm.weight", (H,), "F32", ones=True)
    if not d.tie_embeddings:
        add("output.weight", "lm_head.weight", (V, H), kinds.get("output", "Q6_K"))
    write_gguf(path, meta, tensors)
    return sd
