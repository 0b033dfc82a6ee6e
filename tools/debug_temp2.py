import sys, numpy as np
sys.path.insert(0, '.')
import helix_b200 as hb
from helix_b200 import configs
from helix_b200.engine import CAPTURE_STEP_LOGITS
from oracle import weights
d = configs.tiny_llama(layers=2, head_dim=64, vocab=1000)
sd = weights.llama_state_dict(d, 0, 0.02)
prompt = weights.random_tokens(1, 48, d.vocab)
def drain(e, r):
    out, fin = [], 0
    while not fin:
        e.wait(r, 20000)
        t, fin = e.poll(r)
        out += t
    return out, fin
for trial in range(3):
    with hb.Engine(hb.EngineConfig(max_seqs=4, max_ctx=256, max_batched_tokens=256)) as e:
        e.load_state_dict(d, sd)
        e.start()
        r1 = e.submit(prompt, hb.Sampling(max_tokens=12, eos_token=303))
        r2 = e.submit(prompt, hb.Sampling(max_tokens=200, capture=CAPTURE_STEP_LOGITS))
        print("r1", drain(e, r1))
        e.cancel(r2)
        o2 = drain(e, r2)
        lg = e.captured_logits(r2, CAPTURE_STEP_LOGITS)
        print("r2 n", len(o2[0]), o2[0][:6], "fin", o2[1], "nan rows", np.isnan(lg).any(axis=1).nonzero()[0][:5], e.stats()["cuda_error"])
        a = e.submit(prompt, hb.Sampling(max_tokens=8, temperature=0.8, seed=42, capture=CAPTURE_STEP_LOGITS))
        b = e.submit(prompt, hb.Sampling(max_tokens=8, temperature=0.8, seed=42, capture=CAPTURE_STEP_LOGITS))
        for r in (a, b):
            o = drain(e, r)
            lg = e.captured_logits(r, CAPTURE_STEP_LOGITS)
            print("ab", o, "rows", lg.shape[0], "nan rows", np.isnan(lg).any(axis=1).nonzero()[0][:5], "err", e.stats()["cuda_error"], hb.lib().hb_last_error(e._h))
        e.stop()
