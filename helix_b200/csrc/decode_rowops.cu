// Row kernels of the decode step: each consumes the fp32 partial slabs of a skinny GEMM
// (gemm_skinny.cu), summing them in fixed slab order (bit-deterministic), and fuses the op that follows
// the projection in the decoder layer: RoPE + KV-cache scatter, residual + RMSNorm, SwiGLU, logits.
#include <math.h>

#include "kernels.h"
#include "launch.h"
#include "ptx.cuh"

namespace hb {
namespace {

__device__ __forceinline__ float slab_sum(const float* __restrict__ ws, const uint8_t* __restrict__ segs, int M, int N,
                                          int m, int n) {
  const int ns = segs[n >> 7];
  float acc = 0.f;
  for (int s = 0; s < ns; ++s) acc += ws[((size_t)s * M + m) * N + n];
  return acc;
}
// block-wide entry / exit of a kernel in the decode chain (ptx.cuh dep_*)
__device__ __forceinline__ void block_dep_wait(const int* wait, int count) {
  if (wait) {
    if (threadIdx.x == 0) dep_wait_thread(wait, count);
    __syncthreads();
  } else {
    pdl_wait();
  }
}
__device__ __forceinline__ void block_dep_done(int* done) {
  if (done) {
    __threadfence();   // this thread's stores are visible device-wide ...
    __syncthreads();   // ... for every thread of the block before the block reports
    if (threadIdx.x == 0) dep_signal(done);
  }
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
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.y;
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n < N) out[(size_t)m * ldo + n] = slab_sum(ws, segs, M, N, m, n);
}

// grid (Hq + 2*Hkv, M): one block per (head, token); D/2 threads, one rotary pair each.
__global__ void __launch_bounds__(64)
qkv_rope_kvwrite_kernel(const float* __restrict__ ws, const uint8_t* __restrict__ segs, bf16* __restrict__ qkv_out,
                        const int32_t* __restrict__ positions, const int32_t* __restrict__ slot_mapping,
                        const float* __restrict__ inv_freq, bf16* __restrict__ k_cache, bf16* __restrict__ v_cache,
                        int M, int Hq, int Hkv, int D, int page_size, const int* dep_wait, int dep_count, int* dep_done,
                        const bf16* __restrict__ bias) {
  pdl_launch_dependents();
  block_dep_wait(dep_wait, dep_count);
  const int h = blockIdx.x, m = blockIdx.y, j = threadIdx.x;
  const int half = D / 2;  // == blockDim.x
  const int N = (Hq + 2 * Hkv) * D;
  // GEMM output is rounded to bf16 first (same as the prefill epilogue); RoPE is evaluated in fp32 on top
  const float ba = bias ? __bfloat162float(bias[h * D + j]) : 0.f, bb = bias ? __bfloat162float(bias[h * D + half + j]) : 0.f;
  const float a = bf16_round(slab_sum(ws, segs, M, N, m, h * D + j) + ba);
  const float b = bf16_round(slab_sum(ws, segs, M, N, m, h * D + half + j) + bb);
  const int slot = slot_mapping[m];
  const int page = slot >= 0 ? slot / page_size : 0, off = slot >= 0 ? slot % page_size : 0;
  if (h < Hq + Hkv) {
    float sn, cs;
    sincosf((float)positions[m] * inv_freq[j], &sn, &cs);
    const bf16 lo = __float2bfloat16(rope_lo(a, b, cs, sn)), hi = __float2bfloat16(rope_hi(a, b, cs, sn));
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
  block_dep_done(dep_done);
}

constexpr int kRT = 512;
constexpr int kRMaxVec = 4;  // float4 chunks cached per thread: rows up to 512*4*4 = 8192 wide

__device__ __forceinline__ float4 slab_sum4(const float* __restrict__ ws, const uint8_t* __restrict__ segs, int M, int N,
                                            int m, int n) {
  const int ns = segs[n >> 7];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    if (s < ns) {
      const float4 v = *reinterpret_cast<const float4*>(ws + ((size_t)s * M + m) * N + n);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  for (int s = 8; s < ns; ++s) {
    const float4 v = *reinterpret_cast<const float4*>(ws + ((size_t)s * M + m) * N + n);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  return acc;
}

// one block per row: x = bf16(x + sum of slabs); xn = x * rsqrt(mean(x^2)+eps) * w
__global__ void __launch_bounds__(kRT)
resid_rmsnorm_kernel(const float* __restrict__ ws, const uint8_t* __restrict__ segs, bf16* __restrict__ x,
                     const bf16* __restrict__ w, bf16* __restrict__ xn, int M, int H, float eps, const int* dep_wait,
                     int dep_count, int* dep_done) {
  __shared__ float red[kRT / 32];
  pdl_launch_dependents();
  block_dep_wait(dep_wait, dep_count);
  const int m = blockIdx.x;
  bf16* xr = x + (size_t)m * H;
  float4 cache[kRMaxVec];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < kRMaxVec; ++j) {
    const int n = (threadIdx.x + j * kRT) * 4;
    if (n < H) {
      const float4 a = slab_sum4(ws, segs, M, H, m, n);
      const uint2 xb = *reinterpret_cast<const uint2*>(xr + n);
      const float2 x01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xb.x));
      const float2 x23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xb.y));
      __nv_bfloat162 o01 = __floats2bfloat162_rn(x01.x + a.x, x01.y + a.y);
      __nv_bfloat162 o23 = __floats2bfloat162_rn(x23.x + a.z, x23.y + a.w);
      uint2 ob;
      ob.x = *reinterpret_cast<uint32_t*>(&o01);
      ob.y = *reinterpret_cast<uint32_t*>(&o23);
      *reinterpret_cast<uint2*>(xr + n) = ob;
      const float2 r01 = __bfloat1622float2(o01), r23 = __bfloat1622float2(o23);
      cache[j] = make_float4(r01.x, r01.y, r23.x, r23.y);  // the ROUNDED residual feeds the norm (as in prefill)
      ss += r01.x * r01.x + r01.y * r01.y + r23.x * r23.x + r23.y * r23.y;
    }
  }
  if (!w) {
    block_dep_done(dep_done);
    return;
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float t = (threadIdx.x & 31) < kRT / 32 ? red[threadIdx.x & 31] : 0.f;
  t = warp_sum(t);
  const float inv = rsqrtf(t / (float)H + eps);
#pragma unroll
  for (int j = 0; j < kRMaxVec; ++j) {
    const int n = (threadIdx.x + j * kRT) * 4;
    if (n < H) {
      const uint2 wb = *reinterpret_cast<const uint2*>(w + n);
      const float2 w01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wb.x));
      const float2 w23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wb.y));
      __nv_bfloat162 o01 = __floats2bfloat162_rn(cache[j].x * inv * w01.x, cache[j].y * inv * w01.y);
      __nv_bfloat162 o23 = __floats2bfloat162_rn(cache[j].z * inv * w23.x, cache[j].w * inv * w23.y);
      uint2 ob;
      ob.x = *reinterpret_cast<uint32_t*>(&o01);
      ob.y = *reinterpret_cast<uint32_t*>(&o23);
      *reinterpret_cast<uint2*>(xn + (size_t)m * H + n) = ob;
    }
  }
  block_dep_done(dep_done);
}

__global__ void __launch_bounds__(256)
swiglu_kernel(const float* __restrict__ ws, const uint8_t* __restrict__ segs, bf16* __restrict__ h, int M, int F,
              const int* dep_wait, int dep_count, int* dep_done) {
  pdl_launch_dependents();
  block_dep_wait(dep_wait, dep_count);
  const int m = blockIdx.y;
  const int j = (blockIdx.x * 256 + threadIdx.x) * 4;  // 4 consecutive outputs: never straddles a 128-column tile
  if (j < F) {
  const int t = j >> 7, r = j & 127;
  const float4 g = slab_sum4(ws, segs, M, 2 * F, m, t * 256 + r);
  const float4 u = slab_sum4(ws, segs, M, 2 * F, m, t * 256 + 128 + r);
  __nv_bfloat162 o01 = __floats2bfloat162_rn(g.x / (1.0f + __expf(-g.x)) * u.x, g.y / (1.0f + __expf(-g.y)) * u.y);
  __nv_bfloat162 o23 = __floats2bfloat162_rn(g.z / (1.0f + __expf(-g.z)) * u.z, g.w / (1.0f + __expf(-g.w)) * u.w);
  uint2 ob;
  ob.x = *reinterpret_cast<uint32_t*>(&o01);
  ob.y = *reinterpret_cast<uint32_t*>(&o23);
  *reinterpret_cast<uint2*>(h + (size_t)m * F + j) = ob;
  }
  block_dep_done(dep_done);
}

// one block (128 threads) per row: x = table[token]; xg = bf16(x * gain); ss_out[tile][m] = sum of x^2 over the tile
__global__ void __launch_bounds__(128)
embed_prep_kernel(const int32_t* __restrict__ tokens, const bf16* __restrict__ table, const bf16* __restrict__ gain,
                  bf16* __restrict__ x, bf16* __restrict__ xg, float* __restrict__ ss_out, int H) {
  __shared__ float red[4];
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x, t = threadIdx.x;
  const bf16* src = table + (size_t)tokens[m] * H;
  for (int tile = 0; tile * 128 < H; ++tile) {
    const int n = tile * 128 + t;
    float v = 0.f;
    if (n < H) {
      const bf16 raw = src[n];
      v = __bfloat162float(raw);
      x[(size_t)m * H + n] = raw;
      xg[(size_t)m * H + n] = __float2bfloat16(v * __bfloat162float(gain[n]));
    }
    float sq = warp_sum(v * v);
    __syncthreads();  // red reuse
    if ((t & 31) == 0) red[t >> 5] = sq;
    __syncthreads();
    if (t == 0) ss_out[(size_t)tile * kSkinnySsStride + m] = (red[0] + red[1]) + (red[2] + red[3]);
  }
}

}  // namespace

cudaError_t dec_embed_prep(cudaStream_t s, const int32_t* tokens, const bf16* table, const bf16* gain, bf16* x, bf16* xg,
                           float* ss_out, int M, int H) {
  if (M <= 0) return cudaSuccess;
  return launch_k(embed_prep_kernel, dim3(M), dim3(128), 0, s, true, tokens, table, gain, x, xg, ss_out, H);
}
cudaError_t dec_sum_slabs(cudaStream_t s, const float* ws, const SkinnyPlan& p, float* out, int ldo, int M, int N) {
  return launch_k(sum_slabs_kernel, dim3((N + 255) / 256, M), dim3(256), 0, s, true, ws, p.seg_count, out, ldo, M, N);
}
cudaError_t dec_qkv_rope_kvwrite(cudaStream_t s, const float* ws, const SkinnyPlan& p, bf16* qkv_out,
                                 const int32_t* positions, const int32_t* slot_mapping, const float* inv_freq,
                                 bf16* k_cache, bf16* v_cache, int M, int Hq, int Hkv, int D, int page_size,
                                 const DepSig* dep, const bf16* bias) {
  const DepSig none{};
  if (!dep) dep = &none;
  return launch_k(qkv_rope_kvwrite_kernel, dim3(Hq + 2 * Hkv, M), dim3(D / 2), 0, s, true, ws, p.seg_count, qkv_out,
                  positions, slot_mapping, inv_freq, k_cache, v_cache, M, Hq, Hkv, D, page_size, dep->wait, dep->wait_count,
                  dep->done, bias);
}
int dec_qkv_rope_ctas(int M, int Hq, int Hkv) { return (Hq + 2 * Hkv) * M; }
int dec_swiglu_ctas(int M, int F) { return ((F / 4 + 255) / 256) * M; }
cudaError_t dec_resid_rmsnorm(cudaStream_t s, const float* ws, const SkinnyPlan& p, bf16* x, const bf16* w, bf16* xn,
                              int M, int H, float eps, const DepSig* dep) {
  if (H > kRT * kRMaxVec * 4 || (H % 4)) return cudaErrorInvalidValue;
  const DepSig none{};
  if (!dep) dep = &none;
  return launch_k(resid_rmsnorm_kernel, dim3(M), dim3(kRT), 0, s, true, ws, p.seg_count, x, w, xn, M, H, eps, dep->wait,
                  dep->wait_count, dep->done);
}
cudaError_t dec_swiglu(cudaStream_t s, const float* ws, const SkinnyPlan& p, bf16* h, int M, int F, const DepSig* dep) {
  if (F % 4) return cudaErrorInvalidValue;
  const DepSig none{};
  if (!dep) dep = &none;
  return launch_k(swiglu_kernel, dim3((F / 4 + 255) / 256, M), dim3(256), 0, s, true, ws, p.seg_count, h, M, F, dep->wait,
                  dep->wait_count, dep->done);
}

}  // namespace hb
