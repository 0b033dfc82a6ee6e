"""fp32 numpy restatement of the Llama decoder forward pass (oracle; see oracle/__init__.py).

Follows the model definition the reference's backends implement for `llama` architectures
(llama.cpp `llm_build_llama` / vLLM `LlamaForCausalLM`; HF transformers `modeling_llama.py` is the
pinned comparator): RMSNorm -> q/k/v proj -> RoPE (rotate_half, llama3 frequency scaling) -> causal GQA
attention -> o proj + residual -> RMSNorm -> SwiGLU MLP + residual; final RMSNorm; LM head.
Reference call sites that trigger it: api/pkg/runner/openai_chat_handlers.go:110,140.
"""
import math

import numpy as np


def rope_inv_freq(d):
    dim = d.head_dim
    inv = (1.0 / (np.float32(d.rope_theta) ** (np.arange(0, dim, 2, dtype=np.float32) / np.float32(dim)))).astype(np.float32)
    if getattr(d, "rope_factor", 0.0) and d.rope_factor > 0:
        factor, lo, hi = np.float32(d.rope_factor), np.float32(d.rope_low_freq_factor), np.float32(d.rope_high_freq_factor)
        old = np.float32(d.rope_orig_max_pos)
        low_wl, high_wl = old / lo, old / hi
        wl = (2 * math.pi / inv).astype(np.float32)
        v = np.where(wl > low_wl, inv / factor, inv)
        smooth = (old / wl - lo) / (hi - lo)
        smoothed = (1 - smooth) * v / factor + smooth * v
        medium = ~(wl < high_wl) & ~(wl > low_wl)
        inv = np.where(medium, smoothed, v).astype(np.float32)
    return inv


def rmsnorm(x, w, eps):
    var = np.mean(x.astype(np.float32) ** 2, axis=-1, keepdims=True)
    return (x * (1.0 / np.sqrt(var + np.float32(eps)))) * w


def apply_rope(x, pos, inv_freq):
    """x: [n, heads, D]; pos: [n]. HF rotate_half convention."""
    ang = pos.astype(np.float32)[:, None] * inv_freq[None, :]  # [n, D/2]
    cos = np.concatenate([np.cos(ang), np.cos(ang)], -1)[:, None, :]
    sin = np.concatenate([np.sin(ang), np.sin(ang)], -1)[:, None, :]
    half = x.shape[-1] // 2
    rot = np.concatenate([-x[..., half:], x[..., :half]], -1)
    return x * cos + rot * sin


def silu(x):
    return x / (1.0 + np.exp(-x))


class LlamaOracle:
    def __init__(self, d, sd):
        self.d = d
        self.sd = {k: np.asarray(v, dtype=np.float32) for k, v in sd.items()}
        self.inv_freq = rope_inv_freq(d)
        self.reset()

    def reset(self):
        self.k = [None] * self.d.layers  # [len, Hkv, D]
        self.v = [None] * self.d.layers
        self.len = 0

    def forward(self, tokens):
        """Append `tokens` to the sequence; returns fp32 logits [n, vocab] for the new positions."""
        d, sd = self.d, self.sd
        n = len(tokens)
        pos = np.arange(self.len, self.len + n)
        x = sd["model.embed_tokens.weight"][np.asarray(tokens)]
        g = d.heads // d.kv_heads
        scale = np.float32(1.0 / math.sqrt(d.head_dim))
        for i in range(d.layers):
            p = f"model.layers.{i}."
            xn = rmsnorm(x, sd[p + "input_layernorm.weight"], d.norm_eps)
            q, k, v = (xn @ sd[p + f"self_attn.{t}_proj.weight"].T for t in "qkv")
            if p + "self_attn.q_proj.bias" in sd:  # Qwen2-family decoders (HF modeling_qwen2.py: bias on q/k/v only)
                q, k, v = (y + sd[p + f"self_attn.{t}_proj.bias"] for y, t in zip((q, k, v), "qkv"))
            q = q.reshape(n, d.heads, d.head_dim)
            k = k.reshape(n, d.kv_heads, d.head_dim)
            v = v.reshape(n, d.kv_heads, d.head_dim)
            q = apply_rope(q, pos, self.inv_freq)
            k = apply_rope(k, pos, self.inv_freq)
            self.k[i] = k if self.k[i] is None else np.concatenate([self.k[i], k], 0)
            self.v[i] = v if self.v[i] is None else np.concatenate([self.v[i], v], 0)
            K, V = self.k[i], self.v[i]  # [L, Hkv, D]
            L = K.shape[0]
            out = np.empty((n, d.heads, d.head_dim), np.float32)
            mask = (np.arange(L)[None, :] <= pos[:, None])  # [n, L]
            for h in range(d.heads):
                s = (q[:, h, :] @ K[:, h // g, :].T) * scale
                s = np.where(mask, s, -np.inf)
                s = s - s.max(-1, keepdims=True)
                e = np.exp(s)
                out[:, h, :] = (e / e.sum(-1, keepdims=True)) @ V[:, h // g, :]
            x = x + out.reshape(n, -1) @ sd[p + "self_attn.o_proj.weight"].T
            xn = rmsnorm(x, sd[p + "post_attention_layernorm.weight"], d.norm_eps)
            hmid = silu(xn @ sd[p + "mlp.gate_proj.weight"].T) * (xn @ sd[p + "mlp.up_proj.weight"].T)
            x = x + hmid @ sd[p + "mlp.down_proj.weight"].T
        self.len += n
        x = rmsnorm(x, sd["model.norm.weight"], d.norm_eps)
        self.last_hidden = x.astype(np.float32)  # final-norm hidden states of the new positions (decoder-embedder pooling)
        head = sd["model.embed_tokens.weight"] if d.tie_embeddings else sd["lm_head.weight"]
        return (x @ head.T).astype(np.float32)

    def embed(self, tokens):
        """`--task embed` pooling of a decoder embedder (the reference's default embedding model is one:
        api/pkg/model/models.go:421-433; vLLM pools the LAST token's final hidden state and L2-normalises it)."""
        self.reset()
        self.forward(tokens)
        h = self.last_hidden[-1]
        return h / max(float(np.linalg.norm(h)), 1e-12)

    def greedy(self, prompt, max_tokens):
        """Returns (token ids, logits rows [max_tokens, vocab]) of greedy decoding (argmax, lowest index on ties)."""
        self.reset()
        logits = self.forward(prompt)[-1]
        toks, rows = [], []
        for _ in range(max_tokens):
            rows.append(logits)
            t = int(np.argmax(logits))
            toks.append(t)
            if len(toks) == max_tokens:
                break
            logits = self.forward([t])[-1]
        return toks, np.stack(rows)
