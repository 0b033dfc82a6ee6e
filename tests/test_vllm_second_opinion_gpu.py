"""GPU, opt-in (HB_TEST_VLLM=1): second opinion against the backend the reference actually delegates to
(api/pkg/runner/vllm_runtime.go:163-252 spawns vLLM; this image carries vLLM 0.22, the reference pins 0.11.2).
The seeded Llama-3-8B-shaped checkpoint (2 layers, the HF-pinned fixture's weights and prompt) is written as safetensors
+ config.json, served by vLLM in a subprocess (bf16, greedy, logprobs) and by this engine; greedy token ids must agree
until the first near-tie and the chosen tokens' log-probabilities to 5e-2 (two independent bf16 implementations).
Skipped by default: booting vLLM adds minutes and a second CUDA context to the suite.  Last run: profiles/r02_vllm_parity.txt"""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import helix_b200 as hb
from helix_b200 import configs, weights_io
from oracle import weights

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VLLM_SCRIPT = textwrap.dedent("""
    import json, os, sys
    os.environ.setdefault("VLLM_LOGGING_LEVEL", "WARNING")
    os.environ.setdefault("HF_HUB_OFFLINE", "1")
    from vllm import LLM, SamplingParams
    d, prompt = sys.argv[1], json.loads(sys.argv[2])
    llm = LLM(model=d, skip_tokenizer_init=True, dtype="bfloat16", max_model_len=1024, max_num_seqs=4, gpu_memory_utilization=0.3,
              enforce_eager=True, seed=0, enable_prefix_caching=False)
    sp = SamplingParams(temperature=0.0, max_tokens=8, ignore_eos=True, detokenize=False, logprobs=5)
    out = llm.generate([{"prompt_token_ids": prompt}], sp, use_tqdm=False)[0].outputs[0]
    lps = [{int(k): float(v.logprob) for k, v in step.items()} for step in out.logprobs]
    print("RESULT " + json.dumps({"tokens": [int(t) for t in out.token_ids], "logprobs": lps}))
""")


@pytest.mark.skipif(os.environ.get("HB_TEST_VLLM") != "1", reason="opt-in: HB_TEST_VLLM=1 (boots vLLM in a subprocess)")
def test_greedy_ids_and_logprobs_agree_with_vllm(tmp_path, golden_dir):
    g = np.load(os.path.join(golden_dir, "llama3_8b_2layer.npz"))
    d = configs.llama3_8b()
    d.layers = int(g["layers"])
    sd = weights.llama_state_dict(d, int(g["seed"]), float(g["std"]))
    prompt = g["prompt"].tolist()
    weights_io.write_safetensors(str(tmp_path / "model.safetensors"), sd)
    cfg = {"architectures": ["LlamaForCausalLM"], "model_type": "llama", "hidden_size": d.hidden, "intermediate_size": d.ffn,
           "num_hidden_layers": d.layers, "num_attention_heads": d.heads, "num_key_value_heads": d.kv_heads, "head_dim": d.head_dim,
           "vocab_size": d.vocab, "max_position_embeddings": 8192, "rms_norm_eps": d.norm_eps, "rope_theta": d.rope_theta,
           "torch_dtype": "bfloat16", "tie_word_embeddings": False, "hidden_act": "silu", "bos_token_id": 128000, "eos_token_id": 128001}
    (tmp_path / "config.json").write_text(json.dumps(cfg))
    with hb.Engine(hb.EngineConfig(max_seqs=2, max_ctx=1024, max_batched_tokens=1024, use_cuda_graphs=1)) as e:
        e.load_state_dict(d, sd)
        rids, outs = e.generate([prompt], hb.Sampling(max_tokens=8, logprobs=6))
        ids, lps = e.logprobs(rids[0])
    r = subprocess.run([sys.executable, "-c", VLLM_SCRIPT, str(tmp_path), json.dumps(prompt)], capture_output=True, text=True,
                       timeout=900)
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
    assert line, r.stdout[-2000:] + r.stderr[-4000:]
    v = json.loads(line[0][7:])
    same = 0
    for i, (a, b) in enumerate(zip(outs[0], v["tokens"])):
        if a != b:
            break
        same += 1
        assert abs(float(lps[i, 0]) - v["logprobs"][i][str(a)] if isinstance(next(iter(v["logprobs"][i])), str) else
                   float(lps[i, 0]) - v["logprobs"][i][a]) <= 5e-2
    hf = g["greedy_tokens"].tolist()
    print(f"\n[vLLM second opinion] engine  {outs[0]}\n                      vLLM    {v['tokens']}\n                      HF fp32 {hf}; "
          f"identical for the first {same}/8 steps; engine logprobs {lps[:same, 0].round(4).tolist()}")
    assert same >= 3
    if same < 8:   # after a disagreement: it must be a near-tie in the engine's own distribution
        i = same
        alt = [float(lps[i, k]) for k in range(1, ids.shape[1]) if int(ids[i, k]) == v["tokens"][i]]
        assert alt and float(lps[i, 0]) - alt[0] <= 0.1, (i, outs[0], v["tokens"])
