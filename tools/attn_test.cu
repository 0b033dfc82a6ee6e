// Stand-alone timing harness for the prefill attention kernel (not part of the product library).
//   attn_test [B=8] [S=2048] [Hq=32] [Hkv=8] [D=128] [causal=1] [iters=20]
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../helix_b200/csrc/kernels.h"
using namespace hb;
#ifdef HB_ATTN_TRACE
namespace hb { void attn_trace_dump(); }
#endif
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 8, S = argc > 2 ? atoi(argv[2]) : 2048, Hq = argc > 3 ? atoi(argv[3]) : 32;
  const int Hkv = argc > 4 ? atoi(argv[4]) : 8, D = argc > 5 ? atoi(argv[5]) : 128, causal = argc > 6 ? atoi(argv[6]) : 1;
  const int iters = argc > 7 ? atoi(argv[7]) : 20;
  const int T = B * S, ld = (Hq + 2 * Hkv) * D;
  std::vector<bf16> h((size_t)T * ld);
  uint32_t st = 1;
  for (auto& x : h) { st = st * 1664525u + 1013904223u; x = __float2bfloat16(((st >> 8) & 0xFFFF) / 65536.0f - 0.5f); }
  bf16 *qkv, *out; int32_t* cu;
  CK(cudaMalloc(&qkv, h.size() * 2)); CK(cudaMalloc(&out, (size_t)T * Hq * D * 2)); CK(cudaMalloc(&cu, (B + 1) * 4));
  CK(cudaMemcpy(qkv, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  std::vector<int32_t> hc(B + 1); for (int i = 0; i <= B; ++i) hc[i] = i * S;
  CK(cudaMemcpy(cu, hc.data(), (B + 1) * 4, cudaMemcpyHostToDevice));
  CK(kernels_init());
  AttnPrefillArgs a{};
  a.q = qkv; a.ldq = ld; a.k = qkv + Hq * D; a.ldk = ld; a.v = qkv + (Hq + Hkv) * D; a.ldv = ld; a.out = out; a.ldo = Hq * D;
  a.cu_seqlens = cu; a.B = B; a.T = T; a.max_seqlen = S; a.Hq = Hq; a.Hkv = Hkv; a.D = D; a.causal = causal; a.scale = 1.0f / sqrtf((float)D);
#ifdef HB_ATTN_TRACE
  CK(attn_prefill(0, a));
  CK(cudaDeviceSynchronize());
  attn_trace_dump();
  return 0;
#endif
  for (int i = 0; i < 3; ++i) CK(attn_prefill(0, a));
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) CK(attn_prefill(0, a));
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
  const double flops = 4.0 * B * (double)S * S * D * Hq * (causal ? 0.5 : 1.0);
  printf("attn B=%d S=%d Hq=%d Hkv=%d D=%d causal=%d : %.4f ms  %.1f TFLOP/s\n", B, S, Hq, Hkv, D, causal, ms, flops / (ms * 1e-3) / 1e12);
  return 0;
}
