# full GPU suite, attention ncu capture at the bge shape, bge + llama8b bench lines (no CPU legs)
set -x
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/r02k_pytest_gpu_tail.txt
POLYS=2 bash tools/run_attn.sh 2>&1 | grep "^attn" | tee gpurun_out/r02k_attn_test.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_prefill -c 1 -o gpurun_out/attn_d64_staged -f tools/bin/attn_test 128 512 12 12 64 0 1 > gpurun_out/ncu_attn.log 2>&1
timeout 600 python bench.py --workload bge --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02k_bench_bge.json 2> gpurun_out/r02k_bench_bge.err; tail -c 600 gpurun_out/r02k_bench_bge.json
timeout 600 python bench.py --workload bge --ragged --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02k_bench_bge_ragged.json 2> gpurun_out/r02k_bench_bge_ragged.err; tail -c 300 gpurun_out/r02k_bench_bge_ragged.json
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fixed-total > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; tail -c 300 gpurun_out/r02k_bench.json
