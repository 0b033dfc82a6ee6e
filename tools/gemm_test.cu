// Stand-alone bring-up / timing harness for the tcgen05 GEMM (not part of the product library).
//   gemm_test check           : correctness sweep vs the on-device naive checker
//   gemm_test time M N K epi  : CUDA-event timing
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../helix_b200/csrc/kernels.h"
#include "../helix_b200/csrc/tma_host.h"

using namespace hb;

#define CK(x)                                                                    \
  do {                                                                           \
    cudaError_t e_ = (x);                                                        \
    if (e_ != cudaSuccess) {                                                     \
      printf("CUDA error %s at %s:%d (%s)\n", cudaGetErrorString(e_), __FILE__, __LINE__, tmap_last_error()); \
      exit(2);                                                                   \
    }                                                                            \
  } while (0)

static uint32_t rng_state = 12345;
static float frand() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 65536.0f - 0.5f;
}
static void fill(std::vector<bf16>& v, float scale) {
  for (auto& x : v) x = __float2bfloat16(frand() * scale);
}

static float silu_h(float x) { return x / (1.0f + expf(-x)); }
static float gelu_h(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678f)); }

static int check_one(int M, int N, int K, Epi epi, int block_n) {
  const bool swiglu = epi == EPI_SWIGLU, f32 = epi == EPI_F32;
  const int n_out = swiglu ? N / 2 : N;
  std::vector<bf16> hA((size_t)M * K), hW((size_t)N * K), hR((size_t)M * n_out), hb(N);
  fill(hA, 1.0f); fill(hW, 1.0f); fill(hR, 2.0f); fill(hb, 2.0f);
  bf16 *dA, *dW, *dR, *dbias; void *dC; float* dref;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dW, hW.size() * 2)); CK(cudaMalloc(&dR, hR.size() * 2));
  CK(cudaMalloc(&dbias, N * 2)); CK(cudaMalloc(&dC, (size_t)M * n_out * 4)); CK(cudaMalloc(&dref, (size_t)M * N * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dR, hR.data(), hR.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbias, hb.data(), N * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dC, 0xFF, (size_t)M * n_out * 4));
  GemmArgs g{dA, K, dW, K, dC, n_out, dR, n_out, dbias, M, N, K, epi, block_n};
  GemmArgs gr{dA, K, dW, K, dref, N, nullptr, 0, nullptr, M, N, K, EPI_F32, 0};
  CK(gemm_naive_check(0, gr));
  CK(gemm_bf16_tn(0, g));
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  KERNEL FAILED M=%d N=%d K=%d epi=%d bn=%d: %s\n", M, N, K, epi, block_n, cudaGetErrorString(e)); exit(3); }
  std::vector<float> ref((size_t)M * N);
  CK(cudaMemcpy(ref.data(), dref, ref.size() * 4, cudaMemcpyDeviceToHost));
  std::vector<uint8_t> out((size_t)M * n_out * 4);
  CK(cudaMemcpy(out.data(), dC, out.size(), cudaMemcpyDeviceToHost));
  double max_err = 0; long bad = 0; int shown = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < n_out; ++n) {
      float want;
      if (swiglu) {
        const int tile = n / 128, j = n % 128;
        want = silu_h(ref[(size_t)m * N + tile * 256 + j]) * ref[(size_t)m * N + tile * 256 + 128 + j];
      } else {
        want = ref[(size_t)m * N + n];
        if (epi == EPI_BIAS || epi == EPI_BIAS_GELU || epi == EPI_BIAS_RESID) want += __bfloat162float(hb[n]);
        if (epi == EPI_BIAS_GELU) want = gelu_h(want);
        if (epi == EPI_RESID || epi == EPI_BIAS_RESID) want += __bfloat162float(hR[(size_t)m * n_out + n]);
      }
      float got = f32 ? reinterpret_cast<float*>(out.data())[(size_t)m * n_out + n]
                      : __bfloat162float(reinterpret_cast<bf16*>(out.data())[(size_t)m * n_out + n]);
      float tol = f32 ? 1e-3f + 1e-4f * fabsf(want) : 2e-2f + 1e-2f * fabsf(want);
      float err = fabsf(got - want);
      if (!(err <= tol)) {
        ++bad;
        if (shown < 6) { printf("    mismatch m=%d n=%d got=%f want=%f\n", m, n, got, want); ++shown; }
      }
      if (err > max_err) max_err = err;
    }
  printf("  M=%5d N=%5d K=%5d epi=%d bn=%3d : %s  max_err=%.4g bad=%ld/%ld\n", M, N, K, epi, block_n,
         bad ? "FAIL" : "ok", max_err, bad, (long)M * n_out);
  cudaFree(dA); cudaFree(dW); cudaFree(dR); cudaFree(dbias); cudaFree(dC); cudaFree(dref);
  return bad != 0;
}

static void time_one(int M, int N, int K, Epi epi, int block_n, int iters) {
  const int n_out = epi == EPI_SWIGLU ? N / 2 : N;
  bf16 *dA, *dW, *dR, *dbias; void* dC;
  CK(cudaMalloc(&dA, (size_t)M * K * 2)); CK(cudaMalloc(&dW, (size_t)N * K * 2)); CK(cudaMalloc(&dR, (size_t)M * n_out * 2));
  CK(cudaMalloc(&dbias, N * 2)); CK(cudaMalloc(&dC, (size_t)M * n_out * 4));
  // random-ish non-zero content (power draw depends on data)
  std::vector<bf16> h((size_t)1 << 22); fill(h, 1.0f);
  for (size_t off = 0; off < (size_t)M * K; off += h.size()) CK(cudaMemcpy(dA + off, h.data(), std::min(h.size(), (size_t)M * K - off) * 2, cudaMemcpyHostToDevice));
  for (size_t off = 0; off < (size_t)N * K; off += h.size()) CK(cudaMemcpy(dW + off, h.data(), std::min(h.size(), (size_t)N * K - off) * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dR, 0, (size_t)M * n_out * 2)); CK(cudaMemset(dbias, 0, N * 2));
  GemmArgs g{dA, K, dW, K, dC, n_out, dR, n_out, dbias, M, N, K, epi, block_n};
  for (int i = 0; i < 3; ++i) CK(gemm_bf16_tn(0, g));
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) CK(gemm_bf16_tn(0, g));
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
  double tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12;
  double gb = ((double)M * K + (double)N * K + (double)M * n_out) * 2 / (ms * 1e-3) / 1e9;
  printf("  time M=%5d N=%5d K=%5d epi=%d bn=%3d : %.4f ms  %.1f TFLOP/s  %.0f GB/s(alg)\n", M, N, K, epi, block_n, ms, tf, gb);
  cudaFree(dA); cudaFree(dW); cudaFree(dR); cudaFree(dbias); cudaFree(dC);
}

int main(int argc, char** argv) {
  const char* mode = argc > 1 ? argv[1] : "check";
  if (!strcmp(mode, "check")) {
    int fails = 0;
    // smallest case first: one tile, one k-block
    fails += check_one(128, 64, 64, EPI_F32, 64);
    fails += check_one(128, 64, 64, EPI_NONE, 64);
    fails += check_one(128, 128, 256, EPI_F32, 128);
    fails += check_one(128, 256, 512, EPI_NONE, 256);
    fails += check_one(256, 512, 1024, EPI_NONE, 256);
    fails += check_one(200, 320, 136, EPI_NONE, 64);      // ragged M/N/K tails
    fails += check_one(77, 192, 72, EPI_F32, 128);
    fails += check_one(1024, 2048, 4096, EPI_NONE, 0);     // multi-tile persistent, ring wrap
    fails += check_one(1000, 1536, 768, EPI_BIAS, 0);
    fails += check_one(512, 3072, 768, EPI_BIAS_GELU, 0);
    fails += check_one(384, 768, 3072, EPI_BIAS_RESID, 0);
    fails += check_one(640, 1024, 2048, EPI_RESID, 0);
    fails += check_one(640, 1024, 512, EPI_SWIGLU, 0);
    fails += check_one(33, 1000, 256, EPI_F32, 0);
    fails += check_one(4096, 6144, 4096, EPI_NONE, 0);
    printf("gemm check: %s (%d failing cases)\n", fails ? "FAIL" : "PASS", fails);
    return fails ? 1 : 0;
  }
  if (!strcmp(mode, "time")) {
    if (argc >= 6) {
      time_one(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), (Epi)atoi(argv[5]), argc > 6 ? atoi(argv[6]) : 0, 20);
      return 0;
    }
    // Llama-3-8B prefill shapes at an 8192-token chunk
    for (int bn : {256, 128}) {
      time_one(8192, 6144, 4096, EPI_NONE, bn, 20);
      time_one(8192, 4096, 4096, EPI_RESID, bn, 20);
      time_one(8192, 4096, 14336, EPI_RESID, bn, 10);
    }
    time_one(8192, 28672, 4096, EPI_SWIGLU, 256, 10);
    time_one(8192, 8192, 8192, EPI_NONE, 256, 10);
    // decode-ish
    time_one(32, 6144, 4096, EPI_NONE, 64, 50);
    time_one(32, 28672, 4096, EPI_SWIGLU, 256, 50);
    time_one(32, 4096, 14336, EPI_RESID, 64, 50);
    // BERT
    time_one(32768, 2304, 768, EPI_BIAS, 0, 20);
    time_one(32768, 3072, 768, EPI_BIAS_GELU, 0, 20);
    time_one(32768, 768, 3072, EPI_BIAS_RESID, 0, 20);
    return 0;
  }
  return 0;
}
