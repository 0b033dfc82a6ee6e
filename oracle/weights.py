"""Deterministic random-init checkpoints shared by the oracle, the HF golden generator and the engine.

Every value is bf16-representable (rounded once here), so the fp32 oracle and the bf16 engine start
from bit-identical weights and differ only by activation rounding / accumulation order.
"""
import hashlib

import numpy as np


def to_bf16_f32(a):
    """Round fp32 -> bf16 (nearest even) and return as fp32."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + 0x7FFF
    return (((u + r) >> 16) << 16).astype(np.uint32).view(np.float32)


def _rng(seed, name):
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    return np.random.default_rng(int.from_bytes(h[:8], "little"))


def _t(seed, name, shape, std=0.02, mean=0.0):
    return to_bf16_f32(_rng(seed, name).standard_normal(shape, dtype=np.float32) * std + mean)


def llama_state_dict(d, seed=0, std=0.02):
    """d: helix_b200 ModelDesc-like (hidden, layers, heads, kv_heads, head_dim, ffn, vocab, tie_embeddings)."""
    H, F, V, D = d.hidden, d.ffn, d.vocab, d.head_dim
    sd = {"model.embed_tokens.weight": _t(seed, "embed", (V, H), std)}
    for i in range(d.layers):
        p = f"model.layers.{i}."
        sd[p + "input_layernorm.weight"] = _t(seed, p + "ln1", (H,), 0.05, 1.0)
        sd[p + "self_attn.q_proj.weight"] = _t(seed, p + "q", (d.heads * D, H), std)
        sd[p + "self_attn.k_proj.weight"] = _t(seed, p + "k", (d.kv_heads * D, H), std)
        sd[p + "self_attn.v_proj.weight"] = _t(seed, p + "v", (d.kv_heads * D, H), std)
        if getattr(d, "qkv_bias", 0):
            sd[p + "self_attn.q_proj.bias"] = _t(seed, p + "qb", (d.heads * D,), 0.1)
            sd[p + "self_attn.k_proj.bias"] = _t(seed, p + "kb", (d.kv_heads * D,), 0.1)
            sd[p + "self_attn.v_proj.bias"] = _t(seed, p + "vb", (d.kv_heads * D,), 0.1)
        sd[p + "self_attn.o_proj.weight"] = _t(seed, p + "o", (H, d.heads * D), std)
        sd[p + "post_attention_layernorm.weight"] = _t(seed, p + "ln2", (H,), 0.05, 1.0)
        sd[p + "mlp.gate_proj.weight"] = _t(seed, p + "gate", (F, H), std)
        sd[p + "mlp.up_proj.weight"] = _t(seed, p + "up", (F, H), std)
        sd[p + "mlp.down_proj.weight"] = _t(seed, p + "down", (H, F), std)
    sd["model.norm.weight"] = _t(seed, "norm", (H,), 0.05, 1.0)
    if not d.tie_embeddings:
        sd["lm_head.weight"] = _t(seed, "lm_head", (V, H), std)
    return sd


def bert_state_dict(d, seed=0, std=0.02):
    H, F, V = d.hidden, d.ffn, d.vocab
    tv = d.type_vocab or 2
    sd = {
        "embeddings.word_embeddings.weight": _t(seed, "word", (V, H), std),
        "embeddings.position_embeddings.weight": _t(seed, "pos", (d.max_pos, H), std),
        "embeddings.token_type_embeddings.weight": _t(seed, "type", (tv, H), std),
        "embeddings.LayerNorm.weight": _t(seed, "eln_g", (H,), 0.05, 1.0),
        "embeddings.LayerNorm.bias": _t(seed, "eln_b", (H,), std),
    }
    for i in range(d.layers):
        p = f"encoder.layer.{i}."
        for n in ("query", "key", "value"):
            sd[p + f"attention.self.{n}.weight"] = _t(seed, p + n + "w", (H, H), std)
            sd[p + f"attention.self.{n}.bias"] = _t(seed, p + n + "b", (H,), std)
        sd[p + "attention.output.dense.weight"] = _t(seed, p + "ow", (H, H), std)
        sd[p + "attention.output.dense.bias"] = _t(seed, p + "ob", (H,), std)
        sd[p + "attention.output.LayerNorm.weight"] = _t(seed, p + "ln1g", (H,), 0.05, 1.0)
        sd[p + "attention.output.LayerNorm.bias"] = _t(seed, p + "ln1b", (H,), std)
        sd[p + "intermediate.dense.weight"] = _t(seed, p + "w1", (F, H), std)
        sd[p + "intermediate.dense.bias"] = _t(seed, p + "b1", (F,), std)
        sd[p + "output.dense.weight"] = _t(seed, p + "w2", (H, F), std)
        sd[p + "output.dense.bias"] = _t(seed, p + "b2", (H,), std)
        sd[p + "output.LayerNorm.weight"] = _t(seed, p + "ln2g", (H,), 0.05, 1.0)
        sd[p + "output.LayerNorm.bias"] = _t(seed, p + "ln2b", (H,), std)
    return sd


def random_tokens(seed, n, vocab):
    return _rng(seed, f"tokens{n}").integers(0, vocab, size=n, dtype=np.int64).astype(np.int32)
