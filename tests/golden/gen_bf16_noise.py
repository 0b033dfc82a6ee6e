"""How far does the reference backends' OWN arithmetic (bf16 weights/activations, fp32 accumulate — what vLLM / HF run on a
GPU) sit from the fp32 model, at the BASELINE shapes?  Runs HF transformers on CPU twice (fp32 and bf16) on the seeded
weights / prompts the GPU parity tests use and stores   noise = max_rows |logits_bf16 - logits_fp32|_inf / max(1, |row|_inf)
in tests/golden/bf16_noise.json.  The parity tests state their tolerance next to these numbers: an engine in bf16 cannot be
closer to the fp32 oracle than bf16 rounding allows, and must not be noticeably farther.

    python tests/golden/gen_bf16_noise.py        (minutes; ~12 GB of RAM)
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_golden import hf_llama  # noqa: E402
from helix_b200 import configs  # noqa: E402
from oracle import weights  # noqa: E402

torch.set_grad_enabled(False)


def noise(d, seed, n_prompt):
    sd = weights.llama_state_dict(d, seed, 0.02)
    m = hf_llama(d, sd)
    del sd
    prompt = weights.random_tokens(seed + 1, n_prompt, d.vocab)
    ids = torch.from_numpy(prompt.astype(np.int64))[None]
    ref = m(ids).logits[0].float()
    m = m.to(torch.bfloat16)
    got = m(ids).logits[0].float()
    scale = ref.abs().max(-1).values.clamp_min(1.0)
    per_row = (got - ref).abs().max(-1).values / scale
    flips = int((got.argmax(-1) != ref.argmax(-1)).sum())
    return {"layers": d.layers, "hidden": d.hidden, "seed": seed, "n_prompt": n_prompt, "noise_max": float(per_row.max()),
            "noise_median": float(per_row.median()), "argmax_flips": flips, "torch": torch.__version__}


if __name__ == "__main__":
    out = {}
    d = configs.llama3_8b(); d.layers = 2
    out["llama3_8b_2layer"] = noise(d, 4, 512)
    print(out, flush=True)
    d = configs.llama3_8b(); d.layers = 4
    out["llama3_8b_4layer"] = noise(d, 9, 333)
    print(out, flush=True)
    out["llama32_1b_full"] = noise(configs.llama32_1b(), 0, 128)
    print(out, flush=True)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bf16_noise.json"), "w") as f:
        json.dump(out, f, indent=1)
