// Internal C++ interface to the sm_100a kernels (the C ABI in include/helix_b200.h sits on top).
// All pointers are device pointers; every launcher is asynchronous on `stream` and returns the
// launch status.  Shapes follow the reference backends' conventions (SURVEY.md §2.3, K1–K11).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace hb {

typedef __nv_bfloat16 bf16;

// ---- GEMM: C[M,N] = A[M,K] · W[N,K]^T (both K-major, bf16 in, fp32 accumulate in TMEM) ----
enum Epi : int {
  EPI_NONE = 0,        // C = acc                              (bf16)
  EPI_BIAS = 1,        // C = acc + bias[n]                    (bf16)
  EPI_BIAS_GELU = 2,   // C = gelu_erf(acc + bias[n])          (bf16)
  EPI_RESID = 3,       // C = acc + R[m,n]                     (bf16, R may alias C)
  EPI_BIAS_RESID = 4,  // C = acc + bias[n] + R[m,n]           (bf16)
  EPI_SWIGLU = 5,      // W rows packed per 256-row tile as [128 gate | 128 up];
                       // C[m, tile*128+j] = silu(gate_j) * up_j   (bf16, ldc counts N/2 columns)
  EPI_F32 = 6,         // C = acc                              (fp32)
};

struct GemmArgs {
  const bf16* A;
  int lda;
  const bf16* W;
  int ldw;
  void* C;
  int ldc;
  const bf16* R;
  int ldr;
  const bf16* bias;
  int M, N, K;
  Epi epi;
  int block_n;  // 0 = choose
};
cudaError_t gemm_bf16_tn(cudaStream_t stream, const GemmArgs& a);
// Debug/test-only CUDA-core GEMM (same contract, EPI_NONE/EPI_F32 only); used by tests as an on-device checker.
cudaError_t gemm_naive_check(cudaStream_t stream, const GemmArgs& a);

}  // namespace hb
