# the four encoder GEMM shapes of one bge engine batch (65536 tokens), standalone timing
tools/bin/gemm_test time 65536 2304 768 1 ${BN:-0} | tail -1
tools/bin/gemm_test time 65536 768 768 4 ${BN:-0} | tail -1
tools/bin/gemm_test time 65536 3072 768 2 ${BN:-0} | tail -1
tools/bin/gemm_test time 65536 768 3072 4 ${BN:-0} | tail -1
