# parity tests with the fused RoPE prologue (default), then A/B on the bench
set -x
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_scheduler_gpu.py tests/test_fullsize_gpu.py tests/test_features_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
for f in ${MODES:-1 0 1 0}; do
  HB_DECODE_FUSE_ROPE=$f timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fixed-total > gpurun_out/ab_fr$f.json 2> gpurun_out/ab_fr$f.err
  tail -2 gpurun_out/ab_fr$f.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_fr$f.json"))
print("FUSE_ROPE $f", round(d["value"]), round(d["phases"]["decode_tokens_per_s"]), d["phases"]["decode_ms_per_step"], d["phases"].get("decode_hbm_frac"), d["clocks"]["sm_mhz"], d["gpu_launches"])
PY
done
