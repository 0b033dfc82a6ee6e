// K6: paged-KV decode attention — one query token per sequence, HBM-bound.
// Split-KV: grid (num_splits, Hkv, B); each CTA streams its share of the sequence's KV pages for ONE
// kv head (16-byte cp.async, double buffered, page = contiguous [page_size][D] block per kv head) and
// serves all `G = Hq/Hkv` query heads of the group from the same bytes, so every KV byte is read
// from HBM exactly once per step.  A second tiny kernel merges the per-split (m, l, o) partials.
#include <math.h>

#include "kernels.h"

namespace hb {
namespace {

constexpr int kThreads = 128;
constexpr int kTile = 64;  // kv positions per smem tile (== page_size)
constexpr int kMaxG = 8;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                   static_cast<uint32_t>(__cvta_generic_to_shared(smem))),
               "l"(gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int D, int G>
__global__ void __launch_bounds__(kThreads)
attn_decode_kernel(const bf16* __restrict__ q, int ldq, const bf16* __restrict__ k_cache,
                   const bf16* __restrict__ v_cache, const int32_t* __restrict__ page_table, int max_pages,
                   const int32_t* __restrict__ ctx_lens, float* __restrict__ ws, int Hq, int Hkv, int num_splits,
                   float scale_log2) {
  constexpr int KP = D + 8;  // padded K row (bf16 elements): 16-byte skew kills bank conflicts on row-per-thread reads
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sK = reinterpret_cast<bf16*>(smem_raw);           // [2][kTile][KP]
  bf16* sV = sK + 2 * kTile * KP;                          // [2][kTile][D]
  float* sQ = reinterpret_cast<float*>(sV + 2 * kTile * D);  // [G][D], pre-scaled
  float* sS = sQ + G * D;                                  // [G][kTile]
  float* sF = sS + G * kTile;                              // [G] rescale factors

  const int split = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x;
  const int ctx = ctx_lens[b];
  const int n_pages = (ctx + kTile - 1) / kTile;
  const int pps = (n_pages + num_splits - 1) / num_splits;
  const int p_begin = split * pps;
  const int p_end = min(n_pages, p_begin + pps);
  float* ws_base = ws + ((size_t)(b * Hkv + kvh) * num_splits + split) * G * (D + 2);

  if (p_begin >= p_end) {
    for (int i = tid; i < G * (D + 2); i += kThreads) {
      const int within = i % (D + 2);
      ws_base[i] = (within == D) ? -INFINITY : 0.f;  // m = -inf, l = 0, o = 0
    }
    return;
  }

  for (int i = tid; i < G * D; i += kThreads) {
    const int g = i / D, d = i % D;
    sQ[i] = __bfloat162float(q[(size_t)b * ldq + (kvh * G + g) * D + d]) * scale_log2;
  }

  const int32_t* pt = page_table + (size_t)b * max_pages;
  constexpr int VEC_PER_ROW = D / 8;
  auto load_page = [&](int page_idx, int buf) {
    const size_t base = ((size_t)pt[page_idx] * Hkv + kvh) * kTile * D;
    const uint4* ksrc = reinterpret_cast<const uint4*>(k_cache + base);
    const uint4* vsrc = reinterpret_cast<const uint4*>(v_cache + base);
    bf16* kd = sK + buf * kTile * KP;
    bf16* vd = sV + buf * kTile * D;
    for (int i = tid; i < kTile * VEC_PER_ROW; i += kThreads) {
      const int row = i / VEC_PER_ROW, c = i % VEC_PER_ROW;
      cp_async16(kd + row * KP + c * 8, ksrc + i);
      cp_async16(vd + row * D + c * 8, vsrc + i);
    }
  };

  // per-thread state
  constexpr int NDP = D / 2;                 // d pairs
  constexpr int HSTRIDE = kThreads / NDP;    // head stride in PV phase (2 for D=128, 4 for D=64)
  constexpr int HPT = (G + HSTRIDE - 1) / HSTRIDE;
  float acc[HPT][2];
#pragma unroll
  for (int i = 0; i < HPT; ++i) acc[i][0] = acc[i][1] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;  // meaningful in warp g (lanes replicate)
  const int warp = tid >> 5, lane = tid & 31;

  load_page(p_begin, 0);
  cp_async_commit();
  for (int p = p_begin; p < p_end; ++p) {
    const int buf = (p - p_begin) & 1;
    if (p + 1 < p_end) load_page(p + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const bf16* kb = sK + buf * kTile * KP;
    const bf16* vb = sV + buf * kTile * D;
    const int valid = min(kTile, ctx - p * kTile);

    // ---- scores: thread -> (pos = tid % 64, heads tid/64, tid/64+2, ...)
    {
      const int pos = tid % kTile;
      float s[(G + 1) / 2];
#pragma unroll
      for (int i = 0; i < (G + 1) / 2; ++i) s[i] = 0.f;
      const uint4* krow = reinterpret_cast<const uint4*>(kb + pos * KP);
#pragma unroll 4
      for (int c = 0; c < VEC_PER_ROW; ++c) {
        const uint4 kv = krow[c];
        const __nv_bfloat162* kp = reinterpret_cast<const __nv_bfloat162*>(&kv);
        float kf[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 t = __bfloat1622float2(kp[i]);
          kf[2 * i] = t.x;
          kf[2 * i + 1] = t.y;
        }
#pragma unroll
        for (int i = 0; i < (G + 1) / 2; ++i) {
          const int g = tid / kTile + 2 * i;
          if (g < G) {
            const float4 q0 = *reinterpret_cast<const float4*>(sQ + g * D + c * 8);
            const float4 q1 = *reinterpret_cast<const float4*>(sQ + g * D + c * 8 + 4);
            s[i] += kf[0] * q0.x + kf[1] * q0.y + kf[2] * q0.z + kf[3] * q0.w + kf[4] * q1.x + kf[5] * q1.y +
                    kf[6] * q1.z + kf[7] * q1.w;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < (G + 1) / 2; ++i) {
        const int g = tid / kTile + 2 * i;
        if (g < G) sS[g * kTile + pos] = (pos < valid) ? s[i] : -INFINITY;
      }
    }
    __syncthreads();
    // ---- online softmax: warp g owns head g (extra heads loop)
    for (int g = warp; g < G; g += kThreads / 32) {
      const float s0 = sS[g * kTile + lane], s1 = sS[g * kTile + lane + 32];
      float mx = fmaxf(s0, s1);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      // state for head g lives in registers of warp (g % 4) under index g / 4 — with G <= 4 a single slot
      const float m_new = fmaxf(m_run, mx);
      const float f = (m_run == -INFINITY) ? 0.f : fast_exp2(m_run - m_new);
      const float p0 = fast_exp2(s0 - m_new), p1 = fast_exp2(s1 - m_new);
      float sum = p0 + p1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      l_run = l_run * f + sum;
      m_run = m_new;
      sS[g * kTile + lane] = p0;
      sS[g * kTile + lane + 32] = p1;
      if (lane == 0) sF[g] = f;
    }
    __syncthreads();
    // ---- PV: thread -> (d pair = tid % NDP, heads tid/NDP, +HSTRIDE, ...)
    {
      const int dp = tid % NDP;
#pragma unroll
      for (int i = 0; i < HPT; ++i) {
        const int g = tid / NDP + i * HSTRIDE;
        if (g < G) {
          const float f = sF[g];
          float a0 = acc[i][0] * f, a1 = acc[i][1] * f;
          const float* pr = sS + g * kTile;
#pragma unroll 8
          for (int pos = 0; pos < kTile; ++pos) {
            const float2 vv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(vb + pos * D + dp * 2));
            const float pw = pr[pos];
            a0 += pw * vv.x;
            a1 += pw * vv.y;
          }
          acc[i][0] = a0;
          acc[i][1] = a1;
        }
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();

  // ---- write partials: layout per (b,kvh,split): [G][D+2] = o[D], m, l
  {
    const int dp = tid % NDP;
#pragma unroll
    for (int i = 0; i < HPT; ++i) {
      const int g = tid / NDP + i * HSTRIDE;
      if (g < G) {
        ws_base[g * (D + 2) + dp * 2] = acc[i][0];
        ws_base[g * (D + 2) + dp * 2 + 1] = acc[i][1];
      }
    }
    for (int g = warp; g < G; g += kThreads / 32) {
      if (lane == 0) {
        ws_base[g * (D + 2) + D] = m_run;
        ws_base[g * (D + 2) + D + 1] = l_run;
      }
    }
  }
}

template <int D>
__global__ void __launch_bounds__(D / 2)
attn_decode_combine_kernel(const float* __restrict__ ws, bf16* __restrict__ out, int ldo, int Hq, int Hkv, int G,
                           int num_splits) {
  const int h = blockIdx.x, b = blockIdx.y;
  const int kvh = h / G, g = h % G;
  const float* base = ws + ((size_t)(b * Hkv + kvh) * num_splits) * G * (D + 2) + g * (D + 2);
  const size_t split_stride = (size_t)G * (D + 2);
  float m = -INFINITY;
  for (int s = 0; s < num_splits; ++s) m = fmaxf(m, base[s * split_stride + D]);
  float l = 0.f, o0 = 0.f, o1 = 0.f;
  const int d = threadIdx.x * 2;
  for (int s = 0; s < num_splits; ++s) {
    const float* p = base + s * split_stride;
    const float ms = p[D];
    if (ms == -INFINITY) continue;
    const float w = fast_exp2(ms - m);
    l += w * p[D + 1];
    o0 += w * p[d];
    o1 += w * p[d + 1];
  }
  const float inv = l > 0.f ? 1.0f / l : 0.f;
  *reinterpret_cast<__nv_bfloat162*>(out + (size_t)b * ldo + h * D + d) = __floats2bfloat162_rn(o0 * inv, o1 * inv);
}

template <int D, int G>
constexpr size_t decode_smem() {
  return (size_t)2 * kTile * (D + 8) * 2 + (size_t)2 * kTile * D * 2 + (size_t)G * D * 4 + (size_t)G * kTile * 4 +
         (size_t)kMaxG * 4;
}
template <int D, int G>
cudaError_t set_attr() {
  return cudaFuncSetAttribute(attn_decode_kernel<D, G>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)decode_smem<D, G>());
}

template <int D, int G>
cudaError_t launch(cudaStream_t stream, const AttnDecodeArgs& a) {
  constexpr size_t smem = decode_smem<D, G>();
  auto kern = attn_decode_kernel<D, G>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid(a.num_splits, a.Hkv, a.B);
  kern<<<grid, kThreads, smem, stream>>>(a.q, a.ldq, a.k_cache, a.v_cache, a.page_table, a.max_pages, a.ctx_lens,
                                         a.workspace, a.Hq, a.Hkv, a.num_splits, a.scale * 1.4426950408889634f);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  attn_decode_combine_kernel<D><<<dim3(a.Hq, a.B), D / 2, 0, stream>>>(a.workspace, a.out, a.ldo, a.Hq, a.Hkv, G,
                                                                      a.num_splits);
  return cudaGetLastError();
}

}  // namespace

cudaError_t attn_decode_init() {
  cudaError_t e;
  if ((e = set_attr<128, 1>()) != cudaSuccess) return e;
  if ((e = set_attr<128, 2>()) != cudaSuccess) return e;
  if ((e = set_attr<128, 4>()) != cudaSuccess) return e;
  if ((e = set_attr<64, 1>()) != cudaSuccess) return e;
  if ((e = set_attr<64, 2>()) != cudaSuccess) return e;
  return set_attr<64, 4>();
}

size_t attn_decode_workspace_floats(int B, int Hq, int D, int num_splits) {
  return (size_t)B * Hq * num_splits * (D + 2);
}

cudaError_t attn_decode(cudaStream_t stream, const AttnDecodeArgs& a) {
  if (a.B <= 0) return cudaSuccess;
  if (a.page_size != kTile || a.Hq % a.Hkv || a.num_splits < 1) return cudaErrorInvalidValue;
  const int G = a.Hq / a.Hkv;
  if (a.D == 128) {
    switch (G) {
      case 1: return launch<128, 1>(stream, a);
      case 2: return launch<128, 2>(stream, a);
      case 4: return launch<128, 4>(stream, a);
      default: return cudaErrorInvalidValue;
    }
  }
  if (a.D == 64) {
    switch (G) {
      case 1: return launch<64, 1>(stream, a);
      case 2: return launch<64, 2>(stream, a);
      case 4: return launch<64, 4>(stream, a);
      default: return cudaErrorInvalidValue;
    }
  }
  return cudaErrorInvalidValue;
}

}  // namespace hb
