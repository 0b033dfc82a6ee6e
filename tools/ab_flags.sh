# parity tests with the dependency flags on (default), then A/B on the bench and a timeline
set -x
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_scheduler_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
for f in ${MODES:-1 0}; do
  HB_DECODE_FLAGS=$f timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fixed-total > gpurun_out/ab_flags$f.json 2> gpurun_out/ab_flags$f.err
  tail -3 gpurun_out/ab_flags$f.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_flags$f.json"))
print("FLAGS $f", round(d["value"]), round(d["phases"]["decode_tokens_per_s"]), d["phases"]["decode_ms_per_step"], d["phases"].get("decode_hbm_frac"), d["clocks"]["sm_mhz"], d["latency"]["itl_ms"])
PY
done
HB_DECODE_FLAGS=1 timeout 120 python tools/dec_trace.py 0 > gpurun_out/dec_trace_flags.txt 2>&1; head -16 gpurun_out/dec_trace_flags.txt; tail -3 gpurun_out/dec_trace_flags.txt
