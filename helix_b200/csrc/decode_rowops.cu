// Row kernels of the decode step: each consumes the fp32 partial slabs of a skinny GEMM
// (gemm_skinny.cu), summing them in fixed slab order (bit-deterministic), and fuses the op that follows
// the projection in the decoder layer: RoPE + KV-cache scatter, residual + RMSNorm, SwiGLU, logits.
#include <math.h>

#include "kernels.h"

namespace hb {
namespace {

__device__ __forceinline__ float slab_sum(const float* __restrict__ ws, const uint8_t* __restrict__ segs, int M, int N,
                                          int m, int n) {
  const int ns = segs[n >> 7];
  float acc = 0.f;
  for (int s = 0; s < ns; ++s) acc += ws[((size_t)s * M + m) * N + n];
  return acc;
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256)
sum_slabs_kernel(const float* __restrict__ ws, const uint8_t* __restrict__ segs, float* __restrict__ out, int ldo, int M,
                 int N) {
  const int m = blockIdx.y;
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n < N) out[(size_t)m * ldo + n] = slab_sum(ws, segs, M, N, m, n);
}

// grid (Hq + 2*Hkv, M): one block per (head, token); D/2 threads, one rotary pair each.
__global__ void __launch_bounds__(64)
qkv_rope_kvwrite_kernel(const float* __restrict__ ws, const uint8_t* __restrict__ segs, bf16* __restrict__ qkv_out,
                        const int32_t* __restrict__ positions, const int32_t* __restrict__ slot_mapping,
                        const float* __restrict__ inv_freq, bf16* __restrict__ k_cache, bf16* __restrict__ v_cache,
                        int M, int Hq, int Hkv, int D, int page_size) {
  const int h = blockIdx.x, m = blockIdx.y, j = threadIdx.x;
  const int half = D / 2;
  if (j >= half) return;
  const int N = (Hq + 2 * Hkv) * D;
  // GEMM output is rounded to bf16 first (same as the prefill epilogue); RoPE is evaluated in fp32 on top
  const float a = bf16_round(slab_sum(ws, segs, M, N, m, h * D + j));
  const float b = bf16_round(slab_sum(ws, segs, M, N, m, h * D + half + j));
  const int slot = slot_mapping[m];
  const int page = slot >= 0 ? slot / page_size : 0, off = slot >= 0 ? slot % page_size : 0;
  if (h < Hq + Hkv) {
    float sn, cs;
    sincosf((float)positions[m] * inv_freq[j], &sn, &cs);
    const bf16 lo = __float2bfloat16(a * cs - b * sn), hi = __float2bfloat16(b * cs + a * sn);
    if (h < Hq) {
      bf16* qrow = qkv_out + (size_t)m * N + h * D;
      qrow[j] = lo;
      qrow[half + j] = hi;
    } else if (slot >= 0) {
      const size_t dst = (((size_t)page * Hkv + (h - Hq)) * page_size + off) * D;
      k_cache[dst + j] = lo;
      k_cache[dst + half + j] = hi;
    }
  } else if (slot >= 0) {
    const size_t dst = (((size_t)page * Hkv + (h - Hq - Hkv)) * page_size + off) * D;
    v_cache[dst + j] = __float2bfloat16(a);
    v_cache[dst + half + j] = __float2bfloat16(b);
  }
}

constexpr int kRT = 128;
constexpr int kRC = 8;  // CTAs per row (thread-block cluster): the row's sum of squares is reduced through DSMEM
constexpr int kRMaxPer = 8;  // elements cached per thread: rows up to 8*128*8 = 8192 wide

__device__ __forceinline__ float ld_dsmem_f32(const float* local_ptr, uint32_t cta) {
  uint32_t remote;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;"
               : "=r"(remote)
               : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(local_ptr))), "r"(cta));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
  return v;
}

__global__ void __cluster_dims__(kRC, 1, 1) __launch_bounds__(kRT)
resid_rmsnorm_kernel(const float* __restrict__ ws, const uint8_t* __restrict__ segs, bf16* __restrict__ x,
                     const bf16* __restrict__ w, bf16* __restrict__ xn, int M, int H, float eps) {
  __shared__ float red[kRT / 32];
  __shared__ float part;
  const int m = blockIdx.y;
  const int chunk = (H + kRC - 1) / kRC;
  const int n0 = blockIdx.x * chunk, n1 = min(H, n0 + chunk);
  bf16* xr = x + (size_t)m * H;
  float cache[kRMaxPer];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < kRMaxPer; ++j) {
    const int n = n0 + threadIdx.x + j * kRT;
    if (n < n1) {
      const float v = bf16_round(__bfloat162float(xr[n]) + slab_sum(ws, segs, M, H, m, n));
      cache[j] = v;
      xr[n] = __float2bfloat16(v);
      ss += v * v;
    }
  }
  if (!w) return;  // residual update only (uniform across the whole cluster)
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < kRT / 32; ++i) t += red[i];
    part = t;
  }
  // publish `part` to the cluster, then read all kRC partials in rank order (fixed order -> deterministic)
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  float tot = 0.f;
#pragma unroll
  for (int c = 0; c < kRC; ++c) tot += ld_dsmem_f32(&part, c);
  const float inv = rsqrtf(tot / (float)H + eps);
#pragma unroll
  for (int j = 0; j < kRMaxPer; ++j) {
    const int n = n0 + threadIdx.x + j * kRT;
    if (n < n1) xn[(size_t)m * H + n] = __float2bfloat16(cache[j] * inv * __bfloat162float(w[n]));
  }
  // no CTA may exit while a peer can still read its `part`
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(256)
swiglu_kernel(const float* __restrict__ ws, const uint8_t* __restrict__ segs, bf16* __restrict__ h, int M, int F) {
  const int m = blockIdx.y;
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= F) return;
  const int t = j >> 7, r = j & 127;
  const float g = slab_sum(ws, segs, M, 2 * F, m, t * 256 + r);
  const float u = slab_sum(ws, segs, M, 2 * F, m, t * 256 + 128 + r);
  h[(size_t)m * F + j] = __float2bfloat16(g / (1.0f + __expf(-g)) * u);
}

}  // namespace

cudaError_t dec_sum_slabs(cudaStream_t s, const float* ws, const SkinnyPlan& p, float* out, int ldo, int M, int N) {
  sum_slabs_kernel<<<dim3((N + 255) / 256, M), 256, 0, s>>>(ws, p.seg_count, out, ldo, M, N);
  return cudaGetLastError();
}
cudaError_t dec_qkv_rope_kvwrite(cudaStream_t s, const float* ws, const SkinnyPlan& p, bf16* qkv_out,
                                 const int32_t* positions, const int32_t* slot_mapping, const float* inv_freq,
                                 bf16* k_cache, bf16* v_cache, int M, int Hq, int Hkv, int D, int page_size) {
  qkv_rope_kvwrite_kernel<<<dim3(Hq + 2 * Hkv, M), D / 2, 0, s>>>(ws, p.seg_count, qkv_out, positions, slot_mapping,
                                                                  inv_freq, k_cache, v_cache, M, Hq, Hkv, D, page_size);
  return cudaGetLastError();
}
cudaError_t dec_resid_rmsnorm(cudaStream_t s, const float* ws, const SkinnyPlan& p, bf16* x, const bf16* w, bf16* xn,
                              int M, int H, float eps) {
  if (H > kRC * kRT * kRMaxPer) return cudaErrorInvalidValue;
  resid_rmsnorm_kernel<<<dim3(kRC, M), kRT, 0, s>>>(ws, p.seg_count, x, w, xn, M, H, eps);
  return cudaGetLastError();
}
cudaError_t dec_swiglu(cudaStream_t s, const float* ws, const SkinnyPlan& p, bf16* h, int M, int F) {
  swiglu_kernel<<<dim3((F + 255) / 256, M), 256, 0, s>>>(ws, p.seg_count, h, M, F);
  return cudaGetLastError();
}

}  // namespace hb
