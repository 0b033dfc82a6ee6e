"""The opt-in decode hand-over variants (measured, not faster, kept behind environment switches) and the
row-kernel form of the prefill RoPE (the default does it in the QKV GEMM epilogue) must produce exactly the default
path's tokens.  The switches are read once per process, so every variant runs in its own
interpreter: a seeded tiny Llama, three prompts of different lengths decoded together with CUDA graphs, step logits
captured; token ids must be identical and the logits equal to the last bit (same kernels' arithmetic, different
synchronisation) — except the RoPE prologue, whose slab sums run in a different kernel but in the same order."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import json, sys
import numpy as np
import helix_b200 as hb
from helix_b200 import configs
from helix_b200.engine import CAPTURE_PROMPT_LOGITS, CAPTURE_STEP_LOGITS
from oracle import weights
d = configs.tiny_llama(layers=3, head_dim=128, vocab=1000, rope_scaling=True)
sd = weights.llama_state_dict(d, 7, 0.05)
rng = np.random.default_rng(11)
prompts = [rng.integers(0, 1000, n).astype(np.int32) for n in (5, 70, 131)]
with hb.Engine(hb.EngineConfig(max_seqs=8, max_ctx=512, max_batched_tokens=1024, use_cuda_graphs=1)) as e:
    e.load_state_dict(d, sd)
    rids, outs = e.generate(prompts, hb.Sampling(max_tokens=24, capture=CAPTURE_STEP_LOGITS | CAPTURE_PROMPT_LOGITS))
    logits = [e.captured_logits(r, CAPTURE_STEP_LOGITS) for r in rids]
    plog = [e.captured_logits(r, CAPTURE_PROMPT_LOGITS) for r in rids]
    st = e.stats()
assert st["cuda_error"] == 0 and st["graph_launches"] > 0
print(json.dumps({"tokens": [list(map(int, o)) for o in outs],
                  "digest": [float(np.abs(l).sum()) for l in logits] + [float(np.abs(l).sum()) for l in plog],
                  "last": [l[-1][:8].astype(float).tolist() for l in logits]}))
"""


def run_variant(extra_env):
    env = dict(os.environ)
    for k in ("HB_DECODE_FUSE_ROPE", "HB_DECODE_FLAGS", "HB_DECODE_BANK_MB", "HB_DECODE_FUSED", "HB_PREFILL_FUSE_ROPE"):
        env.pop(k, None)
    env["HB_DECODE_SPLITS"] = "1"  # the variants only engage with one KV split per sequence (the headline batch's case)
    env.update(extra_env)
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    r = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.fixture(scope="module")
def default_run():
    return run_variant({})


@pytest.mark.parametrize("env", [{"HB_DECODE_FUSE_ROPE": "1"}, {"HB_DECODE_FLAGS": "1"}, {"HB_DECODE_BANK_MB": "32"},
                                 {"HB_PREFILL_FUSE_ROPE": "0"}],
                         ids=["rope_prologue", "dependency_flags", "l2_bank", "prefill_rope_row_kernel"])
def test_variant_matches_default(env, default_run):
    got = run_variant(env)
    assert got["tokens"] == default_run["tokens"]
    assert got["last"] == default_run["last"] and got["digest"] == default_run["digest"]
