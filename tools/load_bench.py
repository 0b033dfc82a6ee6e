#!/usr/bin/env python
"""Checkpoint load rate through the C ABI (hb_model_load_begin / hb_model_tensor_set / hb_model_load_finish): writes a
random bf16 safetensors file of the given catalogue shape to a RAM-backed path, then times the load into a fresh engine.
The reference's slot start-up spends "seconds-minutes" here.  One JSON line.
  python tools/load_bench.py [--model meta-llama/Llama-3.2-1B-Instruct] [--dir /dev/shm]"""
import argparse
import json
import os
import struct
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="meta-llama/Llama-3.2-1B-Instruct")
    ap.add_argument("--dir", default="/dev/shm")
    a = ap.parse_args()
    import helix_b200 as hb
    from helix_b200 import weights_io
    from helix_b200.runtime import MODEL_CATALOGUE
    desc = MODEL_CATALOGUE[a.model][0]()
    # tensor names / shapes of the HF checkpoint, contents random bf16 bit patterns (cheap to make: no fp32 detour)
    H, F, V, D = desc.hidden, desc.ffn, desc.vocab, desc.head_dim
    shapes = {"model.embed_tokens.weight": (V, H), "model.norm.weight": (H,)}
    for i in range(desc.layers):
        p = f"model.layers.{i}."
        shapes.update({p + "input_layernorm.weight": (H,), p + "post_attention_layernorm.weight": (H,),
                       p + "self_attn.q_proj.weight": (desc.heads * D, H), p + "self_attn.k_proj.weight": (desc.kv_heads * D, H),
                       p + "self_attn.v_proj.weight": (desc.kv_heads * D, H), p + "self_attn.o_proj.weight": (H, desc.heads * D),
                       p + "mlp.gate_proj.weight": (F, H), p + "mlp.up_proj.weight": (F, H), p + "mlp.down_proj.weight": (H, F)})
    if not desc.tie_embeddings:
        shapes["lm_head.weight"] = (V, H)
    path = os.path.join(a.dir, "hb_load_bench.safetensors")
    rng = np.random.default_rng(0)
    header, off = {}, 0
    for n, s in shapes.items():
        nb = int(np.prod(s)) * 2
        header[n] = {"dtype": "BF16", "shape": list(s), "data_offsets": [off, off + nb]}
        off += nb
    h = json.dumps(header).encode()
    h += b" " * ((8 - len(h) % 8) % 8)
    t0 = time.monotonic()
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(h)))
        f.write(h)
        for n, s in shapes.items():
            # small-magnitude bf16 patterns: exponent byte 0x3c..0x3d
            bits = (rng.integers(0, 1 << 15, int(np.prod(s)), dtype=np.uint16) & 0x81FF) | 0x3C00
            f.write(bits.tobytes())
    t_write = time.monotonic() - t0
    try:
        with hb.Engine(hb.EngineConfig(max_seqs=8, max_ctx=512, max_batched_tokens=1024)) as e:
            t0 = time.monotonic()
            weights_io.load_safetensors(e, desc, path)
            t_load = time.monotonic() - t0
            _, outs = e.generate([np.arange(5, dtype=np.int32)], hb.Sampling(max_tokens=2))
        print(json.dumps({"what": "safetensors -> weight arena through hb_model_tensor_set", "model": a.model, "bytes": off,
                          "write_s": round(t_write, 2), "load_s": round(t_load, 3), "load_gb_per_s": round(off / t_load / 1e9, 2),
                          "source": a.dir + " (page cache / tmpfs)", "generated": [int(t) for t in outs[0]]}))
    finally:
        os.unlink(path)


if __name__ == "__main__":
    main()
