// K5: varlen flash attention for prefill (causal, GQA) and for the BERT encoder
// (bidirectional), on tcgen05 tensor cores with S, P and the output accumulator O all resident in TMEM.
// See the kernel comment for the warp roles and the two-q-tile ping-pong schedule.
#include <math.h>
#include <stdio.h>

#include "kernels.h"
#include "ptx.cuh"
#include "tma_host.h"

namespace hb {
namespace {

constexpr int BQ = 128;     // q rows per softmax warpgroup
constexpr int QPAIR = 256;  // q rows per CTA: two q tiles ping-pong on the tensor pipe
constexpr int BKV = 128;    // kv positions per tile
constexpr int kThreads = 320;
constexpr float kRescaleThreshold = 8.0f;  // log2 units

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int D>
struct ACfg {
  static constexpr int Q_BYTES = BQ * D * 2;    // one q tile
  static constexpr int KV_BYTES = BKV * D * 2;
  static constexpr int SMEM = 2 * Q_BYTES + 4 * KV_BYTES + 1024 + 256;
  static constexpr int SUB = D / 64;  // 64-column swizzle sub-tiles per row
};

// One CTA = one (sequence, q-head, 256-row q pair).  TMEM: S0 | S1 (128 fp32 columns each; P_t, packed bf16x2,
// aliases the first 64 columns of S_t) | O0 | O1.
//   warp 0 lane 0 : TMA producer (Q0,Q1 once; K_j / V_j through 2-stage rings)
//   warp 1 lane 0 : MMA issuer, ping-pong order  PV0(j) S0(j+1) PV1(j) S1(j+1):  while one warpgroup runs its
//                   softmax the tensor pipe works for the other one.  P is consumed straight from TMEM
//                   (tcgen05.mma A-from-TMEM), V straight from its [kv][d] layout (MN-major B): no smem round trip.
//   warps 2..5 / 6..9 : softmax warpgroup of q tile 0 / 1, one thread per q row, two passes over the S row in TMEM
//                   (max, then exp2/sum/pack) so a row never has to live in registers; lazy O rescale.
template <int D>
__global__ void __launch_bounds__(kThreads, 1)
attn_prefill_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, bf16* __restrict__ out, int ldo,
                    const int32_t* __restrict__ cu_seqlens, int group, int causal, float scale_log2,
                    int max_q_pairs) {
  using C = ACfg<D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                        // [2][Q_BYTES]
  uint8_t* sK = sQ + 2 * C::Q_BYTES;         // [2][KV_BYTES]
  uint8_t* sV = sK + 2 * C::KV_BYTES;        // [2][KV_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * C::KV_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2] per q tile
  uint64_t* p_full = bars + 11;   // [2]
  uint64_t* pv_done = bars + 13;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qp = max_q_pairs - 1 - blockIdx.x;  // heaviest (latest) q rows first
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int seq0 = cu_seqlens[b];
  const int len = cu_seqlens[b + 1] - seq0;  // q_len == kv_len (whole-prompt prefill / encoder)
  const int q0 = qp * QPAIR;
  if (q0 >= len) return;
  const int kvh = h / group;
  const int kv_tiles = (len + BKV - 1) / BKV;
  const bool act1 = q0 + BQ < len;
  const int n0 = causal ? min(kv_tiles, q0 / BKV + 1) : kv_tiles;
  const int n1 = act1 ? (causal ? min(kv_tiles, q0 / BKV + 2) : kv_tiles) : 0;
  const int n = max(n0, n1);

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
      mbar_init(&p_full[i], 4);
      mbar_init(&pv_done[i], 1);
    }
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

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, (act1 ? 2 : 1) * C::Q_BYTES);
      for (int t = 0; t < (act1 ? 2 : 1); ++t)
#pragma unroll
        for (int c = 0; c < C::SUB; ++c)
          tma_load_2d(sQ + t * C::Q_BYTES + c * (BQ * 128), &map_q, q_full, h * D + c * 64, seq0 + q0 + t * BQ, kEvictFirst);
      for (int j = 0; j < n; ++j) {
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
      constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, D, 0, 1);  // A = P from TMEM, B = V is MN-major
      auto issue_s = [&](int t, int j) {  // S_t = Q_t · K_j^T
        const uint32_t q_addr = smem_u32(sQ + t * C::Q_BYTES);
        const uint32_t k_addr = smem_u32(sK + (j & 1) * C::KV_BYTES);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t off = (k >> 2) * (128 * 128) + (k & 3) * 32;
          umma_f16_ss(tmem_base + t * BKV, umma_desc_kmajor_sw128(q_addr + off), umma_desc_kmajor_sw128(k_addr + off),
                      idesc_qk, k != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[t]);
      };
      auto issue_pv = [&](int t, int j) {  // O_t += P_t · V_j
        const uint32_t v_addr = smem_u32(sV + (j & 1) * C::KV_BYTES);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          const uint64_t bdesc = umma_desc_mnmajor_sw128(v_addr + k * (16 * 128), BKV * 128, 1024);
          umma_f16_ts(tmem_base + 256 + t * 128, tmem_base + t * BKV + k * 8, bdesc, idesc_pv, (j | k) != 0 ? 1u : 0u);
        }
        umma_commit(&pv_done[t]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      if (n0 > 0) issue_s(0, 0);
      if (n1 > 0) issue_s(1, 0);
      umma_commit(&k_empty[0]);
      for (int j = 0; j < n; ++j) {
        const int s = j & 1;
        mbar_wait(&v_full[s], (j >> 1) & 1);
        if (j < n0) {
          mbar_wait(&p_full[0], j & 1);
          tc_fence_after();
          issue_pv(0, j);
        }
        if (j + 1 < n) {
          mbar_wait(&k_full[s ^ 1], ((j + 1) >> 1) & 1);
          tc_fence_after();
        }
        if (j + 1 < n0) issue_s(0, j + 1);  // overwrites S0/P0 strictly after PV0(j): same-thread MMAs retire in order
        if (j < n1) {
          mbar_wait(&p_full[1], j & 1);
          tc_fence_after();
          issue_pv(1, j);
        }
        umma_commit(&v_empty[s]);
        if (j + 1 < n1) issue_s(1, j + 1);
        if (j + 1 < n) umma_commit(&k_empty[s ^ 1]);
      }
    }
  } else {
    const int t = (warp - 2) >> 2;  // q tile of this warpgroup
    const int nt = t == 0 ? n0 : n1;
    const int qd = warp & 3;
    const int r = qd * 32 + lane;  // q row in tile == TMEM lane
    const int qt0 = q0 + t * BQ;
    const int qpos = qt0 + r;
    const uint32_t lane_sel = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t tS = tmem_base + lane_sel + t * BKV;
    const uint32_t tO = tmem_base + lane_sel + 256 + t * 128;
    float m_used = 0.f, l = 0.f;
    for (int j = 0; j < nt; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      const int kv0 = j * BKV;
      const bool need_mask = (kv0 + BKV > len) || (causal && (kv0 + BKV - 1 > qt0));
      const int lim = causal ? min(len - 1, qpos) : len - 1;  // last valid kv position for this row
      // ---- the S row: four back-to-back TMEM loads, ONE wait (a warpgroup has a single warp per SM sub-partition, so
      //      nothing else hides the load latency), then max / exp2 / pack entirely in registers
      uint32_t sv[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t(&dst)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[c * 32]);
        tmem_ld_32x32b_x32(tS + c * 32, dst);
      }
      tmem_ld_wait();
      if (need_mask) {  // diagonal / ragged tiles only: knock the invalid columns out once, the hot loops stay branch-free
#pragma unroll
        for (int i = 0; i < 128; ++i) sv[i] = (kv0 + i <= lim) ? sv[i] : 0xff800000u;  // -inf
      }
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 128; ++i) mx = fmaxf(mx, __uint_as_float(sv[i]));
      float m_new = (mx == -INFINITY) ? m_used : mx * scale_log2;
      if (j == 0) {
        m_used = m_new;
      } else {
        const bool need = m_new > m_used + kRescaleThreshold;
        mbar_wait(&pv_done[t], (j - 1) & 1);  // O_t consistent (and P_t consumed) before anything below touches them
        tc_fence_after();
        if (__any_sync(0xffffffffu, need)) {
          m_new = fmaxf(m_new, m_used);
          const float f = fast_exp2(m_used - m_new);
          m_used = m_new;
          l *= f;
#pragma unroll
          for (int c = 0; c < D / 32; ++c) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(tO + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st_32x32b_x16(tO + c * 32, *reinterpret_cast<uint32_t(*)[16]>(&o[0]));
            tmem_st_32x32b_x16(tO + c * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&o[16]));
          }
        }
      }
      // ---- p = exp2(s*scale - m) (masked entries are -inf -> 0), row sum, P (bf16x2) written over the S row
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float p0 = fast_exp2(__uint_as_float(sv[c * 16 + 2 * i]) * scale_log2 - m_used);
          const float p1 = fast_exp2(__uint_as_float(sv[c * 16 + 2 * i + 1]) * scale_log2 - m_used);
          sum += p0 + p1;
          pk[i] = pack_bf16x2(p0, p1);
        }
        tmem_st_32x32b_x8(tS + c * 8, pk);
      }
      l += sum;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
    }
    // ---- epilogue: O / l -> global
    if (nt > 0) {
      mbar_wait(&pv_done[t], (nt - 1) & 1);
      tc_fence_after();
      const float inv_l = l > 0.f ? 1.0f / l : 0.f;
      bf16* orow = out + (size_t)(seq0 + qpos) * ldo + h * D;
#pragma unroll
      for (int c = 0; c < D / 32; ++c) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(tO + c * 32, o);
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
  const int max_q_pairs = (a.max_seqlen + QPAIR - 1) / QPAIR;
  dim3 grid(max_q_pairs, a.Hq, a.B);
  const float scale_log2 = a.scale * 1.4426950408889634f;
  kern<<<grid, kThreads, C::SMEM, stream>>>(mq, mk, mv, a.out, a.ldo, a.cu_seqlens, a.Hq / a.Hkv, a.causal, scale_log2,
                                            max_q_pairs);
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
