import sys, numpy as np
sys.path.insert(0, '.')
import helix_b200 as hb
from helix_b200 import configs
from helix_b200.engine import CAPTURE_STEP_LOGITS
from oracle import weights
d = configs.tiny_llama(layers=2, head_dim=64, vocab=1000)
sd = weights.llama_state_dict(d, 0, 0.02)
prompt = weights.random_tokens(1, 48, d.vocab)
for cap in (0, CAPTURE_STEP_LOGITS):
  for temp in (0.0, 0.8):
    with hb.Engine(hb.EngineConfig(max_seqs=4, max_ctx=256, max_batched_tokens=256)) as e:
        e.load_state_dict(d, sd)
        a = e.submit(prompt, hb.Sampling(max_tokens=8, temperature=temp, seed=42, capture=cap))
        e.step()
        b = e.submit(prompt, hb.Sampling(max_tokens=8, temperature=temp, seed=42, capture=cap))
        for i in range(12):
            try:
                e.step()
            except Exception as ex:
                print("step", i, "failed:", ex); break
        print("cap", cap, "temp", temp, e.poll(a), e.poll(b), e.stats()["cuda_error"])
