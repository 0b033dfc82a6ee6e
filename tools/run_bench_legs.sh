# every bench leg once on one B200 (validation run; the driver runs the headline + reference legs itself at round end)
set -x
python bench.py --steps 3 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; tail -3 gpurun_out/r02c_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02c_reference.json 2> gpurun_out/r02c_reference.err; tail -3 gpurun_out/r02c_reference.err
python bench.py --workload bge --steps 1 --warmup 1 > gpurun_out/r02c_bge.json 2> gpurun_out/r02c_bge.err; tail -3 gpurun_out/r02c_bge.err
python bench.py --workload pack --pack-seconds 6 > gpurun_out/r02c_pack.json 2> gpurun_out/r02c_pack.err; tail -5 gpurun_out/r02c_pack.err
python - <<'PY'
import json
for f in ("r02c_bench","r02c_reference","r02c_bge","r02c_pack"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        keys=("value","ms_per_step","e2e","fixed_total","cpu_baseline","phases","modes","solo","latency_spread_arrivals")
        print(f, json.dumps({k:d.get(k) for k in keys if k in d})[:1800])
    except Exception as ex:
        print(f, "FAILED", ex)
PY
