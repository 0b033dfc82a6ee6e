# A/B of the decode step's HBM hand-over bank (HB_DECODE_BANK_MB) on one B200: parity tests first, then short bench runs
set -x
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_features_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
for mb in ${BANKS:-0 32 64}; do
  HB_DECODE_BANK_MB=$mb timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ab_bank$mb.json 2> gpurun_out/ab_bank$mb.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_bank$mb.json"))
print("BANK $mb", round(d["value"]), round(d["phases"]["decode_tokens_per_s"]), round(d["phases"]["decode_ms"],1), d["clocks"]["sm_mhz"], d["latency"]["itl_ms"])
PY
done
