#!/usr/bin/env python
"""Optional same-box comparator: vLLM (the image ships 0.22; the reference pins 0.11.2) on the bench
workload with dummy (random) weights.  Informational only — not the reference arm, not a parity oracle.

    python tools/vllm_compare.py > profiles/r01_vllm_compare.json
"""
import json
import os
import sys
import tempfile
import time

import numpy as np


def main():
    sessions, seq, decode = 32, 2048, 128
    d = tempfile.mkdtemp(prefix="l8b_cfg_")
    cfg = {"architectures": ["LlamaForCausalLM"], "model_type": "llama", "hidden_size": 4096, "intermediate_size": 14336,
           "num_hidden_layers": 32, "num_attention_heads": 32, "num_key_value_heads": 8, "vocab_size": 128256,
           "max_position_embeddings": 8192, "rms_norm_eps": 1e-5, "rope_theta": 500000.0, "torch_dtype": "bfloat16",
           "tie_word_embeddings": False, "hidden_act": "silu", "bos_token_id": 128000, "eos_token_id": 128001}
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump(cfg, f)
    os.environ.setdefault("VLLM_LOGGING_LEVEL", "WARNING")
    os.environ.setdefault("HF_HUB_OFFLINE", "1")
    from vllm import LLM, SamplingParams
    llm = LLM(model=d, load_format="dummy", skip_tokenizer_init=True, dtype="bfloat16", max_model_len=seq + decode + 64,
              max_num_seqs=sessions, max_num_batched_tokens=16384, gpu_memory_utilization=0.5, enforce_eager=False, seed=0,
              enable_prefix_caching=False)  # the timed steps repeat the same prompts: a prefix cache would skip the prefill
    prompts = [{"prompt_token_ids": np.random.default_rng(i).integers(0, 128000, size=seq).tolist()} for i in range(sessions)]
    sp = SamplingParams(temperature=0.0, max_tokens=decode, ignore_eos=True, detokenize=False)
    for _ in range(2):
        llm.generate(prompts, sp, use_tqdm=False)
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        out = llm.generate(prompts, sp, use_tqdm=False)
        times.append(time.perf_counter() - t0)
    n_out = sum(len(o.outputs[0].token_ids) for o in out)
    t = min(times)
    print(json.dumps({"comparator": "vllm", "version": __import__("vllm").__version__, "workload": "Llama-3-8B dummy weights, 32 x (2048 + 128) tokens",
                      "tokens_per_s_e2e": sessions * (seq + decode) / t, "seconds_per_step": t, "generated": n_out,
                      "all_step_seconds": times}))


if __name__ == "__main__":
    main()
