// K5: varlen flash attention for prefill (causal, GQA) and for the BERT encoder
// (bidirectional), on tcgen05 tensor cores with the score tile S and the output accumulator O in TMEM.
//
// One CTA = one (sequence, q-head, 128-row q tile).  Roles:
//   warp 0 lane 0 : TMA producer (Q once; K_j / V_j through 2-stage rings)
//   warp 1 lane 0 : MMA issuer   S_j = Q·K_j^T (SS, both K-major)  and  O += P_j·V_j (A = P from smem,
//                                B = V read MN-major straight from its [kv][d] layout - no transpose pass)
//   warps 2..5    : softmax      one thread per q row: S row TMEM->registers, online max/sum in the
//                                exp2 domain, lazy O rescale (only when the running max moved by > 2^8),
//                                P -> bf16 -> 128B-swizzled smem, final O / l -> global.
// S is double buffered so QK^T of tile j+1 overlaps the softmax of tile j.
#include <math.h>
#include <stdio.h>

#include "kernels.h"
#include "ptx.cuh"
#include "tma_host.h"

namespace hb {
namespace {

constexpr int BQ = 128;   // q rows per CTA
constexpr int BKV = 128;  // kv positions per tile
constexpr int kThreads = 192;
constexpr float kRescaleThreshold = 8.0f;  // log2 units

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int D>
struct ACfg {
  static constexpr int Q_BYTES = BQ * D * 2;
  static constexpr int KV_BYTES = BKV * D * 2;
  static constexpr int P_BYTES = BQ * BKV * 2;
  static constexpr int SMEM = Q_BYTES + 4 * KV_BYTES + P_BYTES + 1024 + 256;
  static constexpr int SUB = D / 64;  // 64-column swizzle sub-tiles per row
};

template <int D>
__global__ void __launch_bounds__(kThreads, 1)
attn_prefill_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, bf16* __restrict__ out, int ldo,
                    const int32_t* __restrict__ cu_seqlens, int group, int causal, float scale_log2,
                    int max_q_tiles) {
  using C = ACfg<D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + C::Q_BYTES;            // [2][KV_BYTES]
  uint8_t* sV = sK + 2 * C::KV_BYTES;       // [2][KV_BYTES]
  uint8_t* sP = sV + 2 * C::KV_BYTES;       // [2 sub-tiles][128 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + C::P_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2]
  uint64_t* s_empty = bars + 11;  // [2]
  uint64_t* p_full = bars + 13;
  uint64_t* pv_done = bars + 14;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = max_q_tiles - 1 - blockIdx.x;  // heaviest (latest) q tiles first
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int seq0 = cu_seqlens[b];
  const int len = cu_seqlens[b + 1] - seq0;  // q_len == kv_len (whole-prompt prefill / encoder)
  const int q0 = qt * BQ;
  if (q0 >= len) return;
  const int kvh = h / group;
  int n_tiles = (len + BKV - 1) / BKV;
  if (causal) n_tiles = min(n_tiles, (q0 + BQ - 1) / BKV + 1);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4);
    }
    mbar_init(p_full, 4);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S = tmem_base;        // 2 x 128 columns
  const uint32_t tmem_O = tmem_base + 256;  // D columns

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, C::Q_BYTES);
#pragma unroll
      for (int c = 0; c < C::SUB; ++c)
        tma_load_2d(sQ + c * (BQ * 128), &map_q, q_full, h * D + c * 64, seq0 + q0, kEvictFirst);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[s], C::KV_BYTES);
#pragma unroll
        for (int c = 0; c < C::SUB; ++c)
          tma_load_2d(sK + s * C::KV_BYTES + c * (BKV * 128), &map_k, &k_full[s], kvh * D + c * 64, seq0 + j * BKV,
                      kEvictLast);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[s], C::KV_BYTES);
#pragma unroll
        for (int c = 0; c < C::SUB; ++c)
          tma_load_2d(sV + s * C::KV_BYTES + c * (BKV * 128), &map_v, &v_full[s], kvh * D + c * 64, seq0 + j * BKV,
                      kEvictLast);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(BQ, BKV, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, D, 0, 1);  // B = V is MN-major
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t p_addr = smem_u32(sP);
      auto issue_qk = [&](int j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_full[s], ph);
        mbar_wait(&s_empty[s], ph ^ 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + s * C::KV_BYTES);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t off = (k >> 2) * (128 * 128) + (k & 3) * 32;
          umma_f16_ss(tmem_S + s * BKV, umma_desc_kmajor_sw128(q_addr + off), umma_desc_kmajor_sw128(k_addr + off),
                      idesc_qk, k != 0 ? 1u : 0u);
        }
        umma_commit(&k_empty[s]);
        umma_commit(&s_full[s]);
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) issue_qk(j + 1);
        const int s = j & 1;
        mbar_wait(p_full, j & 1);
        mbar_wait(&v_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(sV + s * C::KV_BYTES);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          const uint64_t adesc = umma_desc_kmajor_sw128(p_addr + (k >> 2) * (128 * 128) + (k & 3) * 32);
          const uint64_t bdesc = umma_desc_mnmajor_sw128(v_addr + k * (16 * 128), BKV * 128, 1024);
          umma_f16_ss(tmem_O, adesc, bdesc, idesc_pv, (j | k) != 0 ? 1u : 0u);
        }
        umma_commit(&v_empty[s]);
        umma_commit(pv_done);
      }
    }
  } else {
    const int qd = warp & 3;
    const int r = qd * 32 + lane;  // q row in tile == TMEM lane
    const int qpos = q0 + r;
    const uint32_t lane_sel = static_cast<uint32_t>(qd * 32) << 16;
    float m_used = 0.f, l = 0.f;
    uint8_t* p_row = sP + r * 128;
    for (int j = 0; j < n_tiles; ++j) {
      const int s = j & 1;
      mbar_wait(&s_full[s], (j >> 1) & 1);
      tc_fence_after();
      uint32_t sv[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t(&dst)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[c * 32]);
        tmem_ld_32x32b_x32(tmem_S + lane_sel + s * BKV + c * 32, dst);
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[s]);  // S buffer free for QK^T of tile j+2

      const int kv0 = j * BKV;
      const bool need_mask = (kv0 + BKV > len) || (causal && (kv0 + BKV - 1 > q0));
      float mx = -INFINITY;
      if (need_mask) {
        const int lim = causal ? min(len - 1, qpos) : len - 1;  // last valid kv position for this row
#pragma unroll
        for (int i = 0; i < 128; ++i) {
          float x = __uint_as_float(sv[i]);
          x = (kv0 + i <= lim) ? x : -INFINITY;
          sv[i] = __float_as_uint(x);
          mx = fmaxf(mx, x);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 128; ++i) mx = fmaxf(mx, __uint_as_float(sv[i]));
      }
      float m_new = (mx == -INFINITY) ? m_used : mx * scale_log2;
      bool waited_pv = false;
      if (j == 0) {
        m_used = m_new;
      } else {
        const bool need = m_new > m_used + kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          // rescale the running output in TMEM (rare after the first tiles)
          mbar_wait(pv_done, (j - 1) & 1);
          waited_pv = true;
          tc_fence_after();
          m_new = fmaxf(m_new, m_used);
          const float f = exp2f(m_used - m_new);
          m_used = m_new;
          l *= f;
#pragma unroll
          for (int c = 0; c < D / 16; ++c) {
            uint32_t o[16];
            tmem_ld_32x32b_x16(tmem_O + lane_sel + c * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st_32x32b_x16(tmem_O + lane_sel + c * 16, o);
          }
          tmem_st_wait();
        }
      }
      float sum = 0.f;
      uint32_t pk[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const float p0 = fast_exp2(__uint_as_float(sv[2 * i]) * scale_log2 - m_used);
        const float p1 = fast_exp2(__uint_as_float(sv[2 * i + 1]) * scale_log2 - m_used);
        sum += p0 + p1;
        pk[i] = pack_bf16x2(p0, p1);
      }
      l += sum;
      if (j > 0 && !waited_pv) mbar_wait(pv_done, (j - 1) & 1);  // P buffer free (PV of tile j-1 retired)
#pragma unroll
      for (int pc = 0; pc < 16; ++pc) {
        const uint4 o = make_uint4(pk[4 * pc], pk[4 * pc + 1], pk[4 * pc + 2], pk[4 * pc + 3]);
        *reinterpret_cast<uint4*>(p_row + (pc >> 3) * (128 * 128) + (((pc & 7) ^ (r & 7)) << 4)) = o;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> global
    mbar_wait(pv_done, (n_tiles - 1) & 1);
    tc_fence_after();
    const float inv_l = l > 0.f ? 1.0f / l : 0.f;
    bf16* orow = out + (size_t)(seq0 + qpos) * ldo + h * D;
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(tmem_O + lane_sel + c * 32, o);
      tmem_ld_wait();
      if (qpos < len) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[8 * i + 0]) * inv_l, __uint_as_float(o[8 * i + 1]) * inv_l);
          w.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv_l, __uint_as_float(o[8 * i + 3]) * inv_l);
          w.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv_l, __uint_as_float(o[8 * i + 5]) * inv_l);
          w.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv_l, __uint_as_float(o[8 * i + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + c * 32 + i * 8) = w;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int D>
cudaError_t launch(cudaStream_t stream, const AttnPrefillArgs& a) {
  using C = ACfg<D>;
  CUtensorMap mq, mk, mv;
  if (!make_tmap_2d(&mq, a.q, TM_BF16, (uint64_t)a.Hq * D, (uint64_t)a.T, (uint64_t)a.ldq * 2, 64, BQ)) return cudaErrorInvalidValue;
  if (!make_tmap_2d(&mk, a.k, TM_BF16, (uint64_t)a.Hkv * D, (uint64_t)a.T, (uint64_t)a.ldk * 2, 64, BKV)) return cudaErrorInvalidValue;
  if (!make_tmap_2d(&mv, a.v, TM_BF16, (uint64_t)a.Hkv * D, (uint64_t)a.T, (uint64_t)a.ldv * 2, 64, BKV)) return cudaErrorInvalidValue;
  auto kern = attn_prefill_kernel<D>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int max_q_tiles = (a.max_seqlen + BQ - 1) / BQ;
  dim3 grid(max_q_tiles, a.Hq, a.B);
  const float scale_log2 = a.scale * 1.4426950408889634f;
  kern<<<grid, kThreads, C::SMEM, stream>>>(mq, mk, mv, a.out, a.ldo, a.cu_seqlens, a.Hq / a.Hkv, a.causal, scale_log2,
                                            max_q_tiles);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- test-only on-device checker
__global__ void attn_naive_kernel(const bf16* q, int ldq, const bf16* k, int ldk, const bf16* v, int ldv, float* out,
                                  int ldo, const int32_t* cu, int Hq, int group, int D, int causal, float scale) {
  const int b = blockIdx.z, h = blockIdx.y;
  const int seq0 = cu[b], len = cu[b + 1] - seq0;
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= len) return;
  const int kvh = h / group;
  const bf16* qr = q + (size_t)(seq0 + qi) * ldq + h * D;
  float m = -INFINITY, l = 0.f;
  float acc[128];
  for (int d = 0; d < D; ++d) acc[d] = 0.f;
  const int last = causal ? qi : len - 1;
  for (int j = 0; j <= last; ++j) {
    const bf16* kr = k + (size_t)(seq0 + j) * ldk + kvh * D;
    float s = 0.f;
    for (int d = 0; d < D; ++d) s += __bfloat162float(qr[d]) * __bfloat162float(kr[d]);
    s *= scale;
    const float mn = fmaxf(m, s);
    const float f = expf(m - mn), p = expf(s - mn);
    const bf16* vr = v + (size_t)(seq0 + j) * ldv + kvh * D;
    for (int d = 0; d < D; ++d) acc[d] = acc[d] * f + p * __bfloat162float(vr[d]);
    l = l * f + p;
    m = mn;
  }
  for (int d = 0; d < D; ++d) out[(size_t)(seq0 + qi) * ldo + h * D + d] = acc[d] / l;
}

}  // namespace

cudaError_t attn_prefill_init() {
  cudaError_t e = cudaFuncSetAttribute(attn_prefill_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, ACfg<128>::SMEM);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(attn_prefill_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, ACfg<64>::SMEM);
}

cudaError_t attn_prefill(cudaStream_t stream, const AttnPrefillArgs& a) {
  if (a.B <= 0 || a.T <= 0) return cudaSuccess;
  if (a.Hq % a.Hkv) return cudaErrorInvalidValue;
  if ((a.ldq % 8) || (a.ldk % 8) || (a.ldv % 8) || (a.ldo % 8)) return cudaErrorInvalidValue;
  if (a.D == 128) return launch<128>(stream, a);
  if (a.D == 64) return launch<64>(stream, a);
  return cudaErrorInvalidValue;
}

cudaError_t attn_naive_check(cudaStream_t stream, const AttnPrefillArgs& a, float* out_f32) {
  dim3 grid((a.max_seqlen + 63) / 64, a.Hq, a.B);
  attn_naive_kernel<<<grid, 64, 0, stream>>>(a.q, a.ldq, a.k, a.ldk, a.v, a.ldv, out_f32, a.ldo, a.cu_seqlens, a.Hq,
                                             a.Hq / a.Hkv, a.D, a.causal, a.scale);
  return cudaGetLastError();
}

}  // namespace hb
