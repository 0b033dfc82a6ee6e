"""CPU: the numpy oracle must reproduce the HF-transformers golden fixtures (tests/golden/gen_golden.py)."""
import os

import numpy as np
import pytest

from helix_b200 import configs
from oracle import weights
from oracle.bert_ref import bert_embed
from oracle.llama_ref import LlamaOracle, rope_inv_freq

LLAMA_CASES = {
    "llama_tiny_d64": lambda: configs.tiny_llama(layers=2, head_dim=64, vocab=1000),
    "llama_tiny_d64_s05": lambda: configs.tiny_llama(layers=2, head_dim=64, vocab=1000),
    "llama_tiny_d128_rope3": lambda: configs.tiny_llama(layers=3, head_dim=128, vocab=1000, rope_scaling=True),
}


@pytest.mark.parametrize("name", sorted(LLAMA_CASES))
def test_llama_oracle_matches_hf(name, golden_dir):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    d = LLAMA_CASES[name]()
    o = LlamaOracle(d, weights.llama_state_dict(d, int(g["seed"]), float(g["std"])))
    prompt = g["prompt"]
    assert np.array_equal(prompt, weights.random_tokens(int(g["seed"]) + 1, len(prompt), d.vocab))
    logits = o.forward(prompt)
    assert np.abs(logits - g["prompt_logits"]).max() < 2e-4  # fp32 vs fp32, different summation order
    toks, rows = o.greedy(prompt, len(g["greedy_tokens"]))
    assert toks == g["greedy_tokens"].tolist()  # bit-exact token ids
    assert np.abs(rows - g["step_logits"]).max() < 2e-4


def test_rope_llama3_scaling_changes_low_frequencies():
    d = configs.tiny_llama(head_dim=128, rope_scaling=True)
    base = configs.tiny_llama(head_dim=128, rope_scaling=False)
    a, b = rope_inv_freq(d), rope_inv_freq(base)
    assert a.shape == (64,) and np.all(a <= b + 1e-12) and a[-1] < b[-1] / 8 and a[0] == b[0]


def test_bert_oracle_matches_hf(golden_dir):
    g = np.load(os.path.join(golden_dir, "bert_tiny.npz"))
    d = configs.tiny_bert(layers=2, vocab=1000)
    sd = weights.bert_state_dict(d, int(g["seed"]), float(g["std"]))
    lens = g["lens"].tolist()
    offs = np.concatenate([[0], np.cumsum(lens)])
    seqs = [g["tokens"][offs[i]:offs[i + 1]] for i in range(len(lens))]
    e = bert_embed(d, sd, seqs)
    assert np.abs(e - g["embeddings"]).max() < 2e-5
    assert np.allclose(np.linalg.norm(e, axis=1), 1.0, atol=1e-5)


def test_bert_oracle_matches_hf_at_the_full_bge_base_shape(golden_dir):
    """BASELINE configs[2] shape (12 x 768 x 3072, 512 positions): the oracle is pinned to HF there too, not only on the
    tiny config."""
    g = np.load(os.path.join(golden_dir, "bge_base_full.npz"))
    d = configs.bge_base()
    sd = weights.bert_state_dict(d, int(g["seed"]), float(g["std"]))
    lens = g["lens"].tolist()
    offs = np.concatenate([[0], np.cumsum(lens)])
    seqs = [g["tokens"][offs[i]:offs[i + 1]] for i in range(len(lens))]
    e = bert_embed(d, sd, seqs)
    assert np.abs(e - g["embeddings"]).max() < 5e-5


def test_llama_oracle_matches_hf_at_the_llama3_8b_shape(golden_dir):
    """BASELINE configs[1] kernel shapes (hidden 4096, 32/8 heads x 128, ffn 14336, vocab 128256; 2 layers): sampled logit
    rows / columns, every row's argmax and max, and the greedy continuation against HF transformers fp32."""
    g = np.load(os.path.join(golden_dir, "llama3_8b_2layer.npz"))
    d = configs.llama3_8b()
    d.layers = int(g["layers"])
    o = LlamaOracle(d, weights.llama_state_dict(d, int(g["seed"]), float(g["std"])))
    prompt = g["prompt"]
    assert np.array_equal(prompt, weights.random_tokens(int(g["seed"]) + 1, len(prompt), d.vocab))
    logits = o.forward(prompt)
    rows, cols = g["rows"], g["cols"]
    assert np.abs(logits[rows][:, cols] - g["prompt_logits"]).max() < 5e-4
    assert np.abs(logits.max(-1) - g["prompt_max"]).max() < 5e-4
    assert np.array_equal(logits.argmax(-1), g["prompt_argmax"])
    toks, srows = o.greedy(prompt, len(g["greedy_tokens"]))
    assert toks == g["greedy_tokens"].tolist()
    assert np.abs(srows[:, cols] - g["step_logits"]).max() < 5e-4


def test_qwen2_style_decoder_oracle_matches_hf(golden_dir):
    """Qwen2-family decoder (q/k/v biases, GQA group 6, tied embeddings — the architecture of the reference's default
    embedding model, api/pkg/model/models.go:421-433): logits, greedy ids and the last-token embedding vs HF Qwen2ForCausalLM."""
    g = np.load(os.path.join(golden_dir, "qwen2_tiny.npz"))
    d = configs.tiny_qwen2(layers=2, vocab=1000)
    o = LlamaOracle(d, weights.llama_state_dict(d, int(g["seed"]), float(g["std"])))
    prompt = g["prompt"]
    assert np.abs(o.forward(prompt) - g["prompt_logits"]).max() < 2e-4
    toks, rows = o.greedy(prompt, len(g["greedy_tokens"]))
    assert toks == g["greedy_tokens"].tolist() and np.abs(rows - g["step_logits"]).max() < 2e-4
    assert np.abs(o.embed(prompt) - g["embedding"]).max() < 2e-5
