"""CPU: the GGUF reader (hb_gguf_describe / hb_gguf_read_tensor, helix_b200/csrc/gguf.cpp) against the numpy restatement of
ggml's block formats (tests/gguf_ref.py): every supported type dequantises bit for bit, metadata maps to the model
description, llama.cpp's q/k row permutation is undone."""
import numpy as np
import pytest

import helix_b200 as hb
from helix_b200 import configs
from helix_b200.engine import gguf_describe, gguf_read_tensor

from gguf_ref import BLOCK, dequant, hf_permute, llama_gguf, random_blocks, write_gguf


@pytest.mark.parametrize("kind", sorted(BLOCK))
def test_every_block_format_dequantises_bit_exactly(tmp_path, kind):
    rng = np.random.default_rng(hash(kind) % 1000)
    rows, cols = 6, 512
    raw = random_blocks(kind, rows * cols, rng)
    if kind not in ("F32", "F16", "BF16"):   # plus fully random bytes (wild fp16 scales, incl. inf / nan / subnormals)
        wild = rng.integers(0, 256, size=len(raw), dtype=np.uint8).tobytes()
    else:
        wild = raw
    meta = {"general.architecture": "llama", "llama.block_count": 1, "llama.embedding_length": cols, "llama.feed_forward_length": 256,
            "llama.attention.head_count": 4, "llama.attention.head_count_kv": 2}
    path = tmp_path / f"{kind}.gguf"
    write_gguf(path, meta, [("token_embd.weight", kind, (rows, cols), raw), ("output.weight", kind, (rows, cols), wild)])
    for hf_name, blob in (("model.embed_tokens.weight", raw), ("lm_head.weight", wild)):
        got = gguf_read_tensor(path, hf_name)
        want = dequant(kind, blob, rows * cols).reshape(rows, cols)
        assert got.shape == want.shape
        same_bits = np.array_equal(got.view(np.uint32), want.view(np.uint32))
        same_up_to_nan_payload = np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)])
        assert same_bits or same_up_to_nan_payload, kind


def test_metadata_names_and_qk_permutation(tmp_path):
    d = configs.tiny_llama(layers=2, head_dim=64, vocab=1000)
    rng = np.random.default_rng(5)
    path = tmp_path / "tiny.gguf"
    sd = llama_gguf(path, d, rng, {})
    got = gguf_describe(path)
    for f in ("arch", "hidden", "layers", "heads", "kv_heads", "head_dim", "ffn", "vocab", "tie_embeddings", "qkv_bias"):
        assert getattr(got, f) == getattr(d, f), f
    assert abs(got.rope_theta - d.rope_theta) < 1 and abs(got.norm_eps - d.norm_eps) < 1e-9
    for name, want in sd.items():
        w = gguf_read_tensor(path, name)
        assert np.array_equal(w.reshape(want.shape), want), name
    # the permutation helper really is llama.cpp's: permuting the HF-layout truth gives back the rows stored in the file
    q = sd["model.layers.0.self_attn.q_proj.weight"]
    assert not np.array_equal(hf_permute(q, d.heads), q)
    # Qwen2 files carry biases and are NOT permuted
    dq = configs.tiny_qwen2(layers=1, vocab=1000)
    sdq = llama_gguf(tmp_path / "q.gguf", dq, rng, {"attn": "Q8_0"}, arch="qwen2")
    gq = gguf_describe(tmp_path / "q.gguf")
    assert gq.qkv_bias == 1 and gq.tie_embeddings == 1 and gq.heads == 12 and gq.kv_heads == 2
    assert np.array_equal(gguf_read_tensor(tmp_path / "q.gguf", "model.layers.0.self_attn.k_proj.bias").reshape(-1),
                          sdq["model.layers.0.self_attn.k_proj.bias"])


def test_rejected_files(tmp_path):
    (tmp_path / "junk.gguf").write_bytes(b"NOPE" + bytes(64))
    with pytest.raises(hb.HBError):
        gguf_describe(tmp_path / "junk.gguf")
    write_gguf(tmp_path / "moe.gguf", {"general.architecture": "gpt-oss"}, [])
    with pytest.raises(hb.HBError) as e:
        gguf_describe(tmp_path / "moe.gguf")
    assert "not served" in str(e.value)
    meta = {"general.architecture": "llama", "llama.block_count": 1, "llama.embedding_length": 64, "llama.attention.head_count": 1}
    write_gguf(tmp_path / "l31.gguf", meta, [("token_embd.weight", "F32", (8, 64), np.zeros(512, "<f4").tobytes()),
                                             ("rope_freqs.weight", "F32", (32,), np.ones(32, "<f4").tobytes())])
    with pytest.raises(hb.HBError) as e:
        gguf_describe(tmp_path / "l31.gguf")
    assert "rope_freqs" in str(e.value)
