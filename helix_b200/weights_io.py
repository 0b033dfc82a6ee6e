"""Checkpoint I/O for the runtime: safetensors -> hb_model_tensor_set, HF config.json -> ModelDesc.

The reference's backends read HF safetensors (vLLM, HF_HOME cache: api/pkg/runner/vllm_runtime.go:777-812) or GGUF
blobs (Ollama); this runtime takes the HF layout directly.  Pure stdlib + numpy: the header is JSON, tensors are
memory-mapped, bf16 payloads cross the C ABI untouched.
"""
import json
import mmap
import os
import struct

import numpy as np

from .engine import ModelDesc, bf16_bits

_DT = {"BF16": (np.uint16, 2), "F16": (np.float16, 2), "F32": (np.float32, 4)}
_IGNORED = ("pooler.", "embeddings.position_ids", "rotary_emb.inv_freq", "cls.", "visual.")  # visual.: the VL tower (not served)


def read_safetensors(path):
    """Yields (name, dtype string, shape, flat numpy view over the mmap)."""
    f = open(path, "rb")
    n = struct.unpack("<Q", f.read(8))[0]
    header = json.loads(f.read(n))
    mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
    base = 8 + n
    for name, meta in header.items():
        if name == "__metadata__":
            continue
        dt, _ = _DT[meta["dtype"]]
        a, b = meta["data_offsets"]
        yield name, meta["dtype"], tuple(meta["shape"]), np.frombuffer(mm, dtype=dt, count=(b - a) // np.dtype(dt).itemsize, offset=base + a)


def write_safetensors(path, tensors, dtype="BF16"):
    """tensors: {name: fp32 ndarray}; stored as bf16 (or F32). Used by tests and tools."""
    header, blobs, off = {}, [], 0
    for name, a in tensors.items():
        a = np.ascontiguousarray(a, dtype=np.float32)
        raw = bf16_bits(a).tobytes() if dtype == "BF16" else a.tobytes()
        header[name] = {"dtype": dtype, "shape": list(a.shape), "data_offsets": [off, off + len(raw)]}
        blobs.append(raw)
        off += len(raw)
    h = json.dumps(header).encode()
    h += b" " * ((8 - len(h) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(h)))
        f.write(h)
        for b in blobs:
            f.write(b)


def canonical_name(name, arch):
    if arch == 1 and name.startswith("bert."):
        name = name[len("bert."):]
    return name


def to_bf16_bits(dtype, arr):
    if dtype == "BF16":
        return np.ascontiguousarray(arr)
    return bf16_bits(arr.astype(np.float32))


def load_safetensors(engine, desc: ModelDesc, paths):
    """Stream every tensor of the checkpoint shards into the engine's weight arena."""
    import ctypes as C
    l, h = engine._l, engine._h
    engine._ck(l.hb_model_load_begin(h, C.byref(desc.to_c())))
    for path in ([paths] if isinstance(paths, (str, os.PathLike)) else paths):
        for name, dt, shape, arr in read_safetensors(path):
            cname = canonical_name(name, desc.arch)
            if any(cname.startswith(p) or p in cname for p in _IGNORED):
                continue
            if desc.tie_embeddings and cname == "lm_head.weight":
                continue
            bits = to_bf16_bits(dt, arr)
            engine._ck(l.hb_model_tensor_set(h, cname.encode(), bits.ctypes.data, bits.size))
    engine._ck(l.hb_model_load_finish(h))
    engine.desc = desc


def desc_from_hf_config(cfg):
    """HF config.json (dict or path) -> ModelDesc: Llama- and Qwen2-family decoders, BERT-family encoders."""
    if not isinstance(cfg, dict):
        with open(cfg) as f:
            cfg = json.load(f)
    mt = cfg.get("model_type", "")
    if mt in ("qwen2", "qwen2_vl") and "text_config" in cfg:  # VL checkpoints nest the language model's configuration
        cfg = dict(cfg["text_config"], model_type="qwen2")
        mt = "qwen2"
    if mt in ("llama", "qwen2"):
        heads = cfg["num_attention_heads"]
        rope = cfg.get("rope_parameters") or cfg.get("rope_scaling") or {}
        d = ModelDesc(arch=0, hidden=cfg["hidden_size"], layers=cfg["num_hidden_layers"], heads=heads,
                      kv_heads=cfg.get("num_key_value_heads", heads), head_dim=cfg.get("head_dim") or cfg["hidden_size"] // heads,
                      ffn=cfg["intermediate_size"], vocab=cfg["vocab_size"], max_pos=cfg.get("max_position_embeddings", 8192),
                      tie_embeddings=int(bool(cfg.get("tie_word_embeddings", False))), norm_eps=cfg.get("rms_norm_eps", 1e-5),
                      rope_theta=float(rope.get("rope_theta", cfg.get("rope_theta", 10000.0))),
                      qkv_bias=int(mt == "qwen2"))  # Qwen2: biases on q/k/v, otherwise the Llama layer
        if (rope.get("rope_type") or rope.get("type")) == "llama3":
            d.rope_factor = float(rope["factor"])
            d.rope_low_freq_factor = float(rope["low_freq_factor"])
            d.rope_high_freq_factor = float(rope["high_freq_factor"])
            d.rope_orig_max_pos = int(rope["original_max_position_embeddings"])
        return d
    if mt == "bert":
        heads = cfg["num_attention_heads"]
        return ModelDesc(arch=1, hidden=cfg["hidden_size"], layers=cfg["num_hidden_layers"], heads=heads, kv_heads=heads,
                         head_dim=cfg["hidden_size"] // heads, ffn=cfg["intermediate_size"], vocab=cfg["vocab_size"],
                         max_pos=cfg["max_position_embeddings"], type_vocab=cfg.get("type_vocab_size", 2),
                         norm_eps=cfg.get("layer_norm_eps", 1e-12))
    raise ValueError(f"unsupported model_type {mt!r}")
