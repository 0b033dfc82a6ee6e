"""Timeline of one decode step of the headline workload (debug aid): HB_DEC_TRACE=1 makes CTA 0 of every streaming kernel
stamp %globaltimer at fixed points; the dump (one line per kernel, ns relative to the step's first stamp) shows where the
dependency gaps between the kernels are.   python tools/dec_trace.py [bank_mb] > profiles/...
columns: entry | ring issued | hand-over seen | after griddepcontrol.wait | first operands in smem | last load issued |
         last MMA committed | epilogue done"""
import os
import sys

os.environ["HB_DEC_TRACE"] = "1"
os.environ["HB_DEC_TRACE_DUMP"] = "/tmp/dec_trace.txt"
if len(sys.argv) > 1:
    os.environ["HB_DECODE_BANK_MB"] = sys.argv[1]
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import helix_b200 as hb  # noqa: E402
from helix_b200 import configs  # noqa: E402

d = configs.llama3_8b()
e = hb.Engine(hb.EngineConfig(max_seqs=32, max_ctx=2304, max_batched_tokens=16384, use_cuda_graphs=1))
e.load_random(d, 0)
prompts = [np.random.default_rng(i).integers(0, d.vocab, size=2048).astype(np.int32) for i in range(32)]
e.start()
rids = [e.submit(p, hb.Sampling(max_tokens=24)) for p in prompts]
for r in rids:
    fin = 0
    while not fin:
        e.wait(r, 60000)
        _, fin = e.poll(r)
    assert fin == 1, "request failed: " + str(e._l.hb_last_error(e._h))
e.stop()
st = e.stats()
print(f"# decode steps {st['steps_decode']}, gpu_ms_decode {st['gpu_ms_decode']:.2f} -> {st['gpu_ms_decode'] / max(1, st['steps_decode']):.3f} ms/step, bank MB {os.environ.get('HB_DECODE_BANK_MB', '32')}")
rows = [l.split() for l in open("/tmp/dec_trace.txt")]
print("# layer kernel  entry ring_issued handover after_wait first_operands last_load last_mma epilogue_done   (us, relative to layer start)")
layers = sorted({int(r[0]) for r in rows})
first = min(int(x) for r in rows for x in r[2:] if int(x) >= 0)
last = max(int(x) for r in rows for x in r[2:])
print(f"# whole step, first stamp -> last stamp: {(last - first) / 1000:.1f} us; the head row is the LM-head GEMM (then sum / sample kernels follow untraced)")
for l in (0, 1, 15, 31, layers[-1]):
    sel = [r for r in rows if int(r[0]) == l and (l != layers[-1] or r[1] == "head" or l in (0, 1, 15, 31))]
    base = min(int(x) for r in sel for x in r[2:] if int(x) >= 0)
    for r in sel:
        print(f"{r[0]:>3} {r[1]:<5} " + " ".join(f"{(int(x) - base) / 1000:8.2f}" if int(x) >= 0 else "       -" for x in r[2:]))
    print()
e.close()
