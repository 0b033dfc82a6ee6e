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
