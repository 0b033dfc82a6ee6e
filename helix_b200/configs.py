"""Model shapes named by BASELINE.json / SURVEY.md §8a (random-init; no checkpoints offline)."""
from .engine import ModelDesc

LLAMA = 0
BERT = 1


def llama3_8b():
    return ModelDesc(arch=LLAMA, hidden=4096, layers=32, heads=32, kv_heads=8, head_dim=128, ffn=14336, vocab=128256,
                     max_pos=8192, tie_embeddings=0, norm_eps=1e-5, rope_theta=500000.0)


def llama32_1b():
    return ModelDesc(arch=LLAMA, hidden=2048, layers=16, heads=32, kv_heads=8, head_dim=64, ffn=8192, vocab=128256,
                     max_pos=131072, tie_embeddings=1, norm_eps=1e-5, rope_theta=500000.0, rope_factor=32.0,
                     rope_low_freq_factor=1.0, rope_high_freq_factor=4.0, rope_orig_max_pos=8192)


def bge_base():
    return ModelDesc(arch=BERT, hidden=768, layers=12, heads=12, kv_heads=12, head_dim=64, ffn=3072, vocab=30522,
                     max_pos=512, type_vocab=2, norm_eps=1e-12)


def tiny_llama(layers=2, head_dim=64, vocab=1000, rope_scaling=False):
    heads, kv = 4, 2
    return ModelDesc(arch=LLAMA, hidden=heads * head_dim, layers=layers, heads=heads, kv_heads=kv, head_dim=head_dim,
                     ffn=512, vocab=vocab, max_pos=4096, tie_embeddings=0, norm_eps=1e-5, rope_theta=500000.0,
                     rope_factor=32.0 if rope_scaling else 0.0, rope_low_freq_factor=1.0, rope_high_freq_factor=4.0,
                     rope_orig_max_pos=256 if rope_scaling else 0)


def dse_qwen2_2b():
    """Text backbone shape of MrLight/dse-qwen2-2b-mrl-v1 (Qwen2-VL-2B), the reference's default embedding model served
    with `--task embed` (api/pkg/model/models.go:421-433): q/k/v biases, GQA 12/2, tied embeddings, last-token pooling."""
    return ModelDesc(arch=LLAMA, hidden=1536, layers=28, heads=12, kv_heads=2, head_dim=128, ffn=8960, vocab=151936,
                     max_pos=32768, tie_embeddings=1, norm_eps=1e-6, rope_theta=1000000.0, qkv_bias=1)


def tiny_qwen2(layers=2, vocab=1000):
    """Qwen2-style decoder in miniature: q/k/v biases, GQA group 6 (12 q heads over 2 kv heads), tied embeddings."""
    return ModelDesc(arch=LLAMA, hidden=768, layers=layers, heads=12, kv_heads=2, head_dim=64, ffn=512, vocab=vocab,
                     max_pos=4096, tie_embeddings=1, norm_eps=1e-6, rope_theta=1000000.0, qkv_bias=1)


def tiny_bert(layers=2, vocab=1000):
    return ModelDesc(arch=BERT, hidden=256, layers=layers, heads=4, kv_heads=4, head_dim=64, ffn=512, vocab=vocab,
                     max_pos=512, type_vocab=2, norm_eps=1e-12)
