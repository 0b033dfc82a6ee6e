# r02 evidence on one B200: ncu launch list of a shortened bench, full capture of the decode GEMM, sanitizer over the new
# kernels / scheduler tests, vLLM second opinion
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 1 --warmup 1 --layers 4 --decode 16 --no-cpu-baseline --no-fixed-total > gpurun_out/r02_ncu_bench.log 2>&1
tail -2 gpurun_out/r02_ncu_bench.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:gemm_skinny --launch-skip 60 --launch-count 5 -f -o gpurun_out/r02_decode_gemm \
    python bench.py --steps 1 --warmup 1 --layers 2 --decode 8 --no-cpu-baseline --no-fixed-total > gpurun_out/r02_ncu_gemm.log 2>&1
tail -2 gpurun_out/r02_ncu_gemm.log | cut -c1-300
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_scheduler_gpu.py tests/test_features_gpu.py \
    "tests/test_kernels_gpu.py::test_penalties_and_logprobs_kernels_vs_oracle" "tests/test_kernels_gpu.py::test_gemm_skinny_stream_k" \
    "tests/test_engine_gpu.py::test_llama_matches_golden_and_oracle" -m gpu -q -x -p no:cacheprovider > gpurun_out/r02_compute_sanitizer_memcheck.log 2>&1
tail -5 gpurun_out/r02_compute_sanitizer_memcheck.log
HB_TEST_VLLM=1 timeout 1200 python -m pytest tests/test_vllm_second_opinion_gpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r02_vllm_parity.txt 2>&1
tail -15 gpurun_out/r02_vllm_parity.txt
