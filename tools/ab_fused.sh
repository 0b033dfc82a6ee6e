# parity tests on the fused decode path, then A/B: fused finishers vs the separate row kernels (HB_DECODE_FUSED=0)
set -x
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_features_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6
for f in ${MODES:-1 0}; do
  HB_DECODE_FUSED=$f timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fixed-total > gpurun_out/ab_fused$f.json 2> gpurun_out/ab_fused$f.err
  tail -3 gpurun_out/ab_fused$f.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_fused$f.json"))
print("FUSED $f", round(d["value"]), round(d["phases"]["decode_tokens_per_s"]), d["phases"]["decode_ms_per_step"], d["phases"].get("decode_hbm_frac"), d["phases"]["prefill_steps"], d["clocks"]["sm_mhz"], d["latency"]["itl_ms"], d["gpu_launches"])
PY
done
python tools/dec_trace.py 0 > gpurun_out/dec_trace_fused.txt 2>&1; head -12 gpurun_out/dec_trace_fused.txt
