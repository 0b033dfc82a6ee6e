# guard tests for the opt-in decode variants, then memcheck over the attention kernels (staged O write-back, RoPE prologue)
set -x
timeout 900 python -m pytest tests/test_decode_variants_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attn" -p no:cacheprovider > gpurun_out/r02l_memcheck_attn.log 2>&1; echo memcheck rc=$?; tail -4 gpurun_out/r02l_memcheck_attn.log
HB_DECODE_FUSE_ROPE=1 HB_DECODE_SPLITS=1 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_engine_gpu.py -m gpu -q -x -k "llama_tiny_d128 and 1-0" -p no:cacheprovider > gpurun_out/r02l_memcheck_rope.log 2>&1; echo memcheck rc=$?; tail -4 gpurun_out/r02l_memcheck_rope.log
