# the driver's multi-GPU launch line, one N (default 8); JSON line to gpurun_out/
N=${N:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps ${STEPS:-2} --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
echo rc=$?
tail -3 gpurun_out/scale_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/scale_n$N.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","n_gpus","ms_per_step","scaling","gpu_launches")})
print(d.get("fixed_total"))
print(d.get("replica_load"))
PY
