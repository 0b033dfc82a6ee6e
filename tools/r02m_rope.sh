# EPI_ROPE: kernel-level bit-exactness, then the engine suites that run prefill, then A/B on the bench
set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "rope or gemm" -p no:cacheprovider 2>&1 | tail -4
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_scheduler_gpu.py tests/test_features_gpu.py tests/test_baseline_shapes_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
for f in 1 0 1 0; do
  HB_PREFILL_FUSE_ROPE=$f timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fixed-total > gpurun_out/ab_pr$f.json 2> gpurun_out/ab_pr$f.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_pr$f.json"))
print("PREFILL_FUSE_ROPE $f", round(d["value"]), d["phases"]["prefill_ms"], round(d["phases"]["prefill_tokens_per_s"]), d["clocks"]["sm_mhz"], d["gpu_launches"])
PY
done
