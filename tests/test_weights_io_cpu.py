"""CPU: safetensors reader/writer round trip (bit-exact bf16 payloads) and HF config.json -> ModelDesc."""
import numpy as np

from helix_b200 import configs, weights_io
from helix_b200.engine import bf16_bits
from oracle import weights


def test_safetensors_round_trip_bit_exact(tmp_path):
    d = configs.tiny_llama(layers=1, head_dim=64, vocab=200)
    sd = weights.llama_state_dict(d, 3, 0.05)
    p = tmp_path / "m.safetensors"
    weights_io.write_safetensors(p, sd)
    got = {n: (dt, sh, a.copy()) for n, dt, sh, a in weights_io.read_safetensors(p)}
    assert set(got) == set(sd)
    for n, a in sd.items():
        dt, sh, raw = got[n]
        assert dt == "BF16" and sh == a.shape
        assert np.array_equal(raw, bf16_bits(a).ravel())          # bit-exact payload
    p32 = tmp_path / "m32.safetensors"
    weights_io.write_safetensors(p32, {"x": np.arange(6, dtype=np.float32).reshape(2, 3)}, dtype="F32")
    (n, dt, sh, a), = list(weights_io.read_safetensors(p32))
    assert (n, dt, sh) == ("x", "F32", (2, 3)) and np.array_equal(weights_io.to_bf16_bits(dt, a), bf16_bits(np.arange(6)))


def test_desc_from_hf_config_matches_catalogue():
    l8 = weights_io.desc_from_hf_config({
        "model_type": "llama", "hidden_size": 4096, "num_hidden_layers": 32, "num_attention_heads": 32,
        "num_key_value_heads": 8, "intermediate_size": 14336, "vocab_size": 128256, "max_position_embeddings": 8192,
        "rms_norm_eps": 1e-5, "rope_theta": 500000.0, "tie_word_embeddings": False})
    assert l8 == configs.llama3_8b()
    l1 = weights_io.desc_from_hf_config({
        "model_type": "llama", "hidden_size": 2048, "num_hidden_layers": 16, "num_attention_heads": 32, "head_dim": 64,
        "num_key_value_heads": 8, "intermediate_size": 8192, "vocab_size": 128256, "max_position_embeddings": 131072,
        "rms_norm_eps": 1e-5, "tie_word_embeddings": True,
        "rope_scaling": {"rope_type": "llama3", "factor": 32.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                         "original_max_position_embeddings": 8192}, "rope_theta": 500000.0})
    assert l1 == configs.llama32_1b()
    bge = weights_io.desc_from_hf_config({"model_type": "bert", "hidden_size": 768, "num_hidden_layers": 12,
                                          "num_attention_heads": 12, "intermediate_size": 3072, "vocab_size": 30522,
                                          "max_position_embeddings": 512, "type_vocab_size": 2, "layer_norm_eps": 1e-12})
    assert bge == configs.bge_base()
    assert weights_io.canonical_name("bert.encoder.layer.0.output.dense.weight", 1) == "encoder.layer.0.output.dense.weight"


def test_qwen2_config_maps_to_a_biased_llama_layer():
    from helix_b200.weights_io import desc_from_hf_config
    d = desc_from_hf_config({"model_type": "qwen2", "hidden_size": 1536, "num_hidden_layers": 28, "num_attention_heads": 12,
                             "num_key_value_heads": 2, "intermediate_size": 8960, "vocab_size": 151936, "rms_norm_eps": 1e-6,
                             "rope_theta": 1000000.0, "tie_word_embeddings": True, "max_position_embeddings": 32768})
    from helix_b200 import configs
    want = configs.dse_qwen2_2b()
    assert (d.qkv_bias, d.heads, d.kv_heads, d.head_dim, d.hidden, d.ffn, d.vocab, d.tie_embeddings) == \
           (1, want.heads, want.kv_heads, want.head_dim, want.hidden, want.ffn, want.vocab, 1)
    assert desc_from_hf_config({"model_type": "llama", "hidden_size": 256, "num_hidden_layers": 2, "num_attention_heads": 4,
                                "intermediate_size": 512, "vocab_size": 1000}).qkv_bias == 0
