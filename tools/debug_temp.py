import sys, numpy as np
sys.path.insert(0, '.')
import helix_b200 as hb
from helix_b200 import configs
from helix_b200.engine import CAPTURE_STEP_LOGITS
from oracle import weights
d = configs.tiny_llama(layers=2, head_dim=64, vocab=1000)
sd = weights.llama_state_dict(d, 0, 0.02)
prompt = weights.random_tokens(1, 48, d.vocab)
for temp in (0.0, 0.8):
    for nreq in (1, 2):
        with hb.Engine(hb.EngineConfig(max_seqs=4, max_ctx=256, max_batched_tokens=256)) as e:
            e.load_state_dict(d, sd)
            rids = [e.submit(prompt, hb.Sampling(max_tokens=4, temperature=temp, seed=42, capture=CAPTURE_STEP_LOGITS)) for _ in range(nreq)]
            for _ in range(8):
                try:
                    e.step()
                except Exception as ex:
                    print("step failed", ex); break
            for r in rids:
                t, fin = e.poll(r)
                lg = e.captured_logits(r, CAPTURE_STEP_LOGITS)
                print(f"temp={temp} nreq={nreq} toks={t} fin={fin} logits rows={lg.shape[0]} nan={np.isnan(lg).sum()} absmax={np.nanmax(np.abs(lg)) if lg.size else None}")
