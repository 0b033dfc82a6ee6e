#!/bin/bash
# builds tools/bin/{gemm_test,attn_test} from the library sources (bring-up / timing harnesses)
set -e
cd "$(dirname "$0")/../helix_b200/csrc"
mkdir -p ../../tools/bin
SRC="gemm.cu tma_host.cpp attn_prefill.cu attn_decode.cu gemm_skinny.cu"
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17"
nvcc $F -o ../../tools/bin/gemm_test ../../tools/gemm_test.cu $SRC &
nvcc $F -o ../../tools/bin/attn_test ../../tools/attn_test.cu $SRC &
wait
