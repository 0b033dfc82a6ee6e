# final state of the round: POLY sweep of the attention harness, full GPU suite, smoke, default bench line, bge line
set -x
for p in 0 2 4; do HB_ATTN_POLY=$p tools/bin/attn_test 128 512 12 12 64 0 20; HB_ATTN_POLY=$p tools/bin/attn_test 8 2048 32 8 128 1 20; done 2>&1 | grep "^attn" | tee gpurun_out/r02n_attn_poly.txt
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/r02n_pytest_gpu_tail.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; tail -c 400 gpurun_out/r02n_bench.json
timeout 600 python bench.py --workload bge --steps 1 --warmup 1 > gpurun_out/r02n_bench_bge.json 2> gpurun_out/r02n_bench_bge.err; tail -c 300 gpurun_out/r02n_bench_bge.json
