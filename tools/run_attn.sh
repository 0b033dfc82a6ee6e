# attention kernel: parity tests, then the timing harness at the bge shape / long sequences / llama shape
set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attn or attention" -p no:cacheprovider 2>&1 | tail -3
for p in ${POLYS:-0 2 4}; do HB_ATTN_POLY=$p tools/bin/attn_test 128 512 12 12 64 0 20; done
tools/bin/attn_test 16 4096 12 12 64 0 10
tools/bin/attn_test 128 512 12 12 128 0 20
tools/bin/attn_test 8 2048 32 8 128 1 20
tools/bin/attn_test 2 16384 32 8 128 1 5
