"""fp32 numpy restatement of the BERT encoder + CLS pooling + L2 normalisation (oracle).

Model definition behind the reference's `/v1/embeddings` proxy (api/pkg/runner/openai_embedding_handlers.go:569;
vLLM `--task embed`, api/pkg/model/models.go:421-436); HF transformers `modeling_bert.py` is the pinned
comparator: embeddings (word + position + token_type 0) -> LayerNorm -> N x [self-attention with biases ->
dense + residual -> LayerNorm -> dense + GELU(erf) -> dense + residual -> LayerNorm] -> CLS row -> x/||x||.
"""
import math

import numpy as np

try:
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)


def layernorm(x, g, b, eps):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * g + b


def gelu(x):
    return (0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))).astype(np.float32)


def bert_hidden(d, sd, tokens):
    """Last hidden state [n, hidden] for ONE sequence (no padding)."""
    n = len(tokens)
    x = sd["embeddings.word_embeddings.weight"][np.asarray(tokens)] + sd["embeddings.position_embeddings.weight"][:n] + \
        sd["embeddings.token_type_embeddings.weight"][0][None, :]
    x = layernorm(x, sd["embeddings.LayerNorm.weight"], sd["embeddings.LayerNorm.bias"], d.norm_eps)
    scale = 1.0 / math.sqrt(d.head_dim)
    for i in range(d.layers):
        p = f"encoder.layer.{i}."
        q = (x @ sd[p + "attention.self.query.weight"].T + sd[p + "attention.self.query.bias"]).reshape(n, d.heads, -1)
        k = (x @ sd[p + "attention.self.key.weight"].T + sd[p + "attention.self.key.bias"]).reshape(n, d.heads, -1)
        v = (x @ sd[p + "attention.self.value.weight"].T + sd[p + "attention.self.value.bias"]).reshape(n, d.heads, -1)
        ctx = np.empty_like(q)
        for h in range(d.heads):
            s = (q[:, h] @ k[:, h].T) * scale
            s = s - s.max(-1, keepdims=True)
            e = np.exp(s)
            ctx[:, h] = (e / e.sum(-1, keepdims=True)) @ v[:, h]
        a = ctx.reshape(n, -1) @ sd[p + "attention.output.dense.weight"].T + sd[p + "attention.output.dense.bias"]
        x = layernorm(a + x, sd[p + "attention.output.LayerNorm.weight"], sd[p + "attention.output.LayerNorm.bias"], d.norm_eps)
        m = gelu(x @ sd[p + "intermediate.dense.weight"].T + sd[p + "intermediate.dense.bias"])
        o = m @ sd[p + "output.dense.weight"].T + sd[p + "output.dense.bias"]
        x = layernorm(o + x, sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"], d.norm_eps)
    return x.astype(np.float32)


def bert_embed(d, sd, seqs):
    """CLS-pooled, L2-normalised embeddings [len(seqs), hidden] (bge convention)."""
    sd = {k: np.asarray(v, dtype=np.float32) for k, v in sd.items()}
    out = np.empty((len(seqs), d.hidden), np.float32)
    for i, s in enumerate(seqs):
        cls = bert_hidden(d, sd, s)[0]
        out[i] = cls / max(float(np.linalg.norm(cls)), 1e-12)
    return out
