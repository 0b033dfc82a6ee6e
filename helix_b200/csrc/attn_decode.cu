// K6: paged-KV decode attention — one query token per sequence, HBM-bound —
// on the tcgen05 tensor cores with TRANSPOSED score / output tiles so that all 128 TMEM lanes work:
//
//   S^T[kv=128 lanes, 16 cols] = K_tile[128 kv x D] (A, K-major smem)  ·  Q^T   (B: the G<=16 query heads of
//                                                                               the GQA group, K-major)
//   O^T[d =128 lanes, 16 cols] += V_tile^T (A, read MN-major straight from the [kv][d] page layout) · P^T (B)
//
// Split-KV grid (num_splits, Hkv, B): each CTA streams its share of the sequence's pages for ONE kv head with
// TMA (3-D tensor map over [page*Hkv][64 tok][D], 8 KB boxes; K_j and V_j tiles alternate through one 3-slot ring,
// ~106 KB per CTA, so TWO CTAs share an SM and one CTA's prologue / epilogue hides behind the other's stream) and
// serves all G query heads from the same bytes: every KV byte is read from HBM exactly once per step.
//   warp 0 lane 0 : TMA producer      warp 1 lane 0 : MMA issuer
//   warps 2..5    : softmax — thread <-> kv position (4 fp32 scores each): tile max / sum by warp shuffles +
//                   one smem hop, lazy rescale of O^T (only when the running max moved by > 2^8), P^T -> smem.
// A second tiny kernel merges the per-split (m, l, o) partials.
#include <math.h>
#include <stdio.h>

#include <algorithm>

#include "kernels.h"
#include "launch.h"
#include "ptx.cuh"
#include "tma_host.h"

namespace hb {
namespace {

constexpr int kThreads = 192;
constexpr int BKV = 128;   // kv positions per tile (2 pages)
constexpr int PAGE = 64;
constexpr int NQ = 16;     // UMMA N: query heads of the group, zero padded
constexpr int kMaxG = 16;
constexpr float kRescaleThreshold = 8.0f;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int D>
struct DCfg {
  static constexpr int STAGES = 3;                   // ring slots; K_j and V_j tiles alternate through ONE ring
  static constexpr int KV_BYTES = BKV * D * 2;       // one K (or V) tile
  static constexpr int Q_BYTES = NQ * D * 2;
  static constexpr int P_BYTES = NQ * BKV * 2;
  static constexpr int TAIL = (D == 64) ? BKV * 128 : 0;  // the M=128 PV MMA of a D=64 head reads one tile-chunk past V
  static constexpr int SMEM = STAGES * KV_BYTES + TAIL + Q_BYTES + P_BYTES + 1024 + 1024;  // ~106 KB at D=128: 2 CTAs/SM
  static constexpr int SUB = D / 64;
};

// MN-major A/B operand descriptor and instruction descriptor with A MN-major are shared with attn_prefill (ptx.cuh).

template <int D, int G>
__global__ void __launch_bounds__(kThreads, 2)
attn_decode_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                   const bf16* __restrict__ q, int ldq, const int32_t* __restrict__ page_table, int max_pages,
                   const int32_t* __restrict__ ctx_lens, float* __restrict__ ws, int Hkv, int num_splits,
                   float scale_log2, bf16* __restrict__ out_direct, int ldo, const int* sig_wait, int sig_wait_count,
                   int* sig_done, int bank_tiles, const bf16* __restrict__ k_cache, const bf16* __restrict__ v_cache,
                   unsigned long long* trace, const int* dep_wait, int dep_count, int* dep_done,
                   const __grid_constant__ DecodeRope rope) {
  using C = DCfg<D>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;                                 // [STAGES][KV_BYTES] (+TAIL): items K_0 V_0 K_1 V_1 ...
  uint8_t* sQ = ring + STAGES * C::KV_BYTES + C::TAIL;  // [SUB][16 rows][128 B]
  uint8_t* sP = sQ + C::Q_BYTES;                        // [2][16 rows][128 B]
  float* sRed = reinterpret_cast<float*>(sP + C::P_BYTES);  // [2][4 warps][G]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRed + 2 * 4 * kMaxG);
  uint64_t* r_full = bars;                  // [STAGES]
  uint64_t* r_empty = bars + STAGES;        // [STAGES]
  uint64_t* s_full = bars + 2 * STAGES;     // [2]
  uint64_t* s_empty = bars + 2 * STAGES + 2;  // [2]
  uint64_t* q_ready = bars + 2 * STAGES + 4;
  uint64_t* p_full = bars + 2 * STAGES + 5;
  uint64_t* pv_done = bars + 2 * STAGES + 6;
  uint64_t* kv_ready = bars + 2 * STAGES + 7;  // fused RoPE prologue: this CTA's new K/V row is in the cache
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 8);

  // Programmatic dependent launch: only two things here depend on the kernel before this one (RoPE + KV write of the
  // step's new token): q, and the LAST kv tile (it holds the new position).  Everything else — page table, context
  // lengths (uploaded before the step), all older K/V pages — is immutable during the step, so the TMA ring is filled and
  // the barriers / TMEM are set up while the predecessor is still draining; griddepcontrol.wait sits right in front of
  // the first dependent access of each role.
  pdl_launch_dependents();
  if (blockIdx.x | blockIdx.y | blockIdx.z) trace = nullptr;
  if (threadIdx.x == 0) trace_ev(trace, 0);
  const int split = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ctx = ctx_lens[b];
  const int n_pages = (ctx + PAGE - 1) / PAGE;
  const int n_tiles_total = (ctx + BKV - 1) / BKV;
  const int tps = (n_tiles_total + num_splits - 1) / num_splits;
  const int t_begin = split * tps;
  const int t_end = min(n_tiles_total, t_begin + tps);
  float* ws_base = ws + ((size_t)(b * Hkv + kvh) * num_splits + split) * G * (D + 2);
  if (t_begin >= t_end) {
    if (dep_wait) {
      if (threadIdx.x == 0) dep_wait_thread(dep_wait, dep_count);
      __syncthreads();
    } else {
      pdl_wait();  // ws may still be read by an earlier launch's combine pass
    }
    for (int i = threadIdx.x; i < G * (D + 2); i += kThreads) ws_base[i] = (i % (D + 2) == D) ? -INFINITY : 0.f;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {  // every CTA reports, also one without work
      sig_add(sig_done);
      dep_signal(dep_done);
    }
    return;
  }
  const int n_tiles = t_end - t_begin;
  const int32_t* pt = page_table + (size_t)b * max_pages;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&r_full[i], 1);
      mbar_init(&r_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4);
    }
    mbar_init(q_ready, 4);
    mbar_init(p_full, 4);
    mbar_init(pv_done, 1);
    mbar_init(kv_ready, 4);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);  // warp-uniform for the compiler too
  const uint32_t tmem_S = tmem_base;       // 2 x 16 columns
  const uint32_t tmem_O = tmem_base + 32;  // 16 columns

  // Single-thread roles: whole warp converged through loops and waits, TMA / tcgen05 instructions predicated on an
  // elect.sync leader (under `if (lane == 0)` ptxas wraps each of them in a waterfall loop; see gemm.cu).
  if (warp == 0) {
    const bool elected = elect_one_sync();
    {
      for (int j = 0; j < n_tiles; ++j) {
        const int pg0 = (t_begin + j) * 2;
        if (t_begin + j == n_tiles_total - 1) {  // the tile with the step's new position
          if (rope.ws) {  // written by this CTA's own softmax warps (fused prologue), fenced for the async proxy
            mbar_wait(kv_ready, 0);
          } else if (dep_wait) {
            if (elected) dep_wait_thread(dep_wait, dep_count);
            __syncwarp();
            if (elected) fence_proxy_async_all();
          } else {
            pdl_wait();
          }
        }
        // a tile's second page may not exist yet: re-load the first one (finite data, masked by the softmax)
        const int page_a = pt[pg0], page_b = pt[min(pg0 + 1, n_pages - 1)];
#pragma unroll
        for (int kv = 0; kv < 2; ++kv) {  // item 2j = K_j, item 2j+1 = V_j
          const int item = 2 * j + kv;
          const int s = item % STAGES;
          mbar_wait(&r_empty[s], ((item / STAGES) & 1) ^ 1);
          const CUtensorMap* mp = kv ? &map_v : &map_k;
          uint8_t* dst = ring + s * C::KV_BYTES;
          if (elected) {
            mbar_arrive_expect_tx(&r_full[s], C::KV_BYTES);
#pragma unroll
            for (int c = 0; c < C::SUB; ++c) {
              tma_load_3d(dst + c * (BKV * 128), mp, &r_full[s], c * 64, 0, page_a * Hkv + kvh, kEvictFirst);
              tma_load_3d(dst + c * (BKV * 128) + PAGE * 128, mp, &r_full[s], c * 64, 0, page_b * Hkv + kvh, kEvictFirst);
            }
          }
        }
      }
      if (elected) { sig_add(sig_done); trace_ev(trace, 5); }  // this CTA's HBM demand ends here
    }
  } else if (warp == 1) {
    const bool elected = elect_one_sync();
    {
      constexpr uint32_t idesc_s = umma_idesc_bf16(BKV, NQ, 0, 0);   // A = K tile (K-major), B = Q (K-major)
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, NQ, 1, 0);   // A = V tile (MN-major: d contiguous), B = P^T
      const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
      auto issue_s = [&](int j) {
        const int item = 2 * j, s = item % STAGES;
        mbar_wait(&r_full[s], (item / STAGES) & 1);
        mbar_wait(&s_empty[j & 1], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(ring + s * C::KV_BYTES);
        if (elected) {
#pragma unroll
          for (int k = 0; k < D / 16; ++k) {
            const uint64_t adesc = umma_desc_kmajor_sw128(k_addr + (k >> 2) * (BKV * 128) + (k & 3) * 32);
            const uint64_t bdesc = umma_desc_kmajor_sw128(q_addr + (k >> 2) * (NQ * 128) + (k & 3) * 32);
            umma_f16_ss(tmem_S + (j & 1) * NQ, adesc, bdesc, idesc_s, k != 0 ? 1u : 0u);
          }
          umma_commit(&r_empty[s]);
          umma_commit(&s_full[j & 1]);
        }
      };
      mbar_wait(q_ready, 0);
      issue_s(0);
      if (elected) trace_ev(trace, 4);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) issue_s(j + 1);
        const int item = 2 * j + 1, s = item % STAGES;
        mbar_wait(p_full, j & 1);
        mbar_wait(&r_full[s], (item / STAGES) & 1);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(ring + s * C::KV_BYTES);
        if (elected) {
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k) {
            const uint64_t adesc = umma_desc_mnmajor_sw128(v_addr + k * (16 * 128), BKV * 128, 1024);
            const uint64_t bdesc = umma_desc_kmajor_sw128(p_addr + (k >> 2) * (NQ * 128) + (k & 3) * 32);
            umma_f16_ss(tmem_O, adesc, bdesc, idesc_o, (j | k) != 0 ? 1u : 0u);
          }
          umma_commit(&r_empty[s]);
          umma_commit(pv_done);
        }
      }
    }
  } else {
    const int qd = warp & 3;
    const int t = qd * 32 + lane;  // kv position within the tile (S^T lane) / head-dim index (O^T lane)
    const uint32_t lane_sel = static_cast<uint32_t>(qd * 32) << 16;
    // HBM hand-over (ptx.cuh sig_*): tiles 0 and 1 are already on their way into the ring.  Once the previous streaming
    // kernel has issued its last load, these (still idle) warps pull the next `bank_tiles` K/V tiles into L2, one 128-byte
    // line of each page per thread, so HBM keeps working while this kernel waits for its q.
    if (bank_tiles > 0 && n_tiles > 2) {
      if (lane == 0) sig_wait_ge(sig_wait, sig_wait_count);
      __syncwarp();
      constexpr int LINES = PAGE * D * 2 / 128;  // 128-byte lines per (page, kv head) block
      const int j_end = min(n_tiles, 2 + bank_tiles);
      for (int jj = 2; jj < j_end; ++jj) {
        const int p0 = (t_begin + jj) * 2;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int pg = pt[min(p0 + h, n_pages - 1)];
          const size_t off = ((size_t)pg * Hkv + kvh) * PAGE * D + (size_t)t * 64;
          if (t < LINES) {
            prefetch_l2_line(k_cache + off);
            prefetch_l2_line(v_cache + off);
          }
        }
      }
    }
    // ---- stage Q^T (G rows of D, rows G..15 zero) into the K-major swizzled B-operand layout; zero P
    if (t == 0) trace_ev(trace, 2);
    if (dep_wait) {  // q is written by the predecessor; so is nothing else this role reads
      if (t == 0) dep_wait_thread(dep_wait, dep_count);
      bar_sync(1, 128);
    } else {
      pdl_wait();
    }
    if (t == 0) trace_ev(trace, 3);
    if (rope.ws) {
      // fused prologue: q heads of this group, and the sequence's new K / V row, from the QKV GEMM's slabs
      constexpr int HALF = D / 2;
      const int Hq = Hkv * G;
      const int slot = rope.slots[b], pos = rope.positions[b];
      const int page = slot >= 0 ? slot / PAGE : 0, off = slot >= 0 ? slot % PAGE : 0;
      for (int i = t; i < C::Q_BYTES / 16; i += 128) reinterpret_cast<uint4*>(sQ)[i] = make_uint4(0, 0, 0, 0);  // rows G..15 stay zero
      for (int i = t; i < C::P_BYTES / 16; i += 128) reinterpret_cast<uint4*>(sP)[i] = make_uint4(0, 0, 0, 0);
      bar_sync(1, 128);
      auto slab = [&](int n) {  // fixed slab order, as in decode_rowops.cu
        const int ns = rope.segs[n >> 7];
        float acc = 0.f;
        for (int sgi = 0; sgi < ns; ++sgi) acc += rope.ws[((size_t)sgi * rope.M + b) * rope.N + n];
        return acc;
      };
      auto q_elem = [&](int g, int e) {  // address of element e of q row g in the K-major swizzled B-operand layout
        const int c = e >> 3;
        return reinterpret_cast<bf16*>(sQ + (c >> 3) * (NQ * 128) + g * 128 + (((c & 7) ^ (g & 7)) << 4) + (e & 7) * 2);
      };
      for (int pp = t; pp < (G + 2) * HALF; pp += 128) {
        const int hl = pp / HALF, j = pp % HALF;
        const int head = hl < G ? kvh * G + hl : (hl == G ? Hq + kvh : Hq + Hkv + kvh);
        const int n0 = head * D + j, n1 = n0 + HALF;
        const float b0 = rope.bias ? __bfloat162float(rope.bias[n0]) : 0.f, b1 = rope.bias ? __bfloat162float(rope.bias[n1]) : 0.f;
        // the projection output is rounded to bf16 first (as in the prefill epilogue); RoPE is evaluated in fp32 on top
        const float a = __bfloat162float(__float2bfloat16(slab(n0) + b0)), bb = __bfloat162float(__float2bfloat16(slab(n1) + b1));
        if (hl <= G) {
          float sn, cs;
          sincosf((float)pos * rope.inv_freq[j], &sn, &cs);
          const bf16 lo = __float2bfloat16(rope_lo(a, bb, cs, sn)), hi = __float2bfloat16(rope_hi(a, bb, cs, sn));
          if (hl < G) {
            *q_elem(hl, j) = lo;
            *q_elem(hl, HALF + j) = hi;
          } else if (slot >= 0) {
            const size_t dst = (((size_t)page * Hkv + kvh) * PAGE + off) * D;
            rope.k_cache[dst + j] = lo;
            rope.k_cache[dst + HALF + j] = hi;
          }
        } else if (slot >= 0) {
          const size_t dst = (((size_t)page * Hkv + kvh) * PAGE + off) * D;
          rope.v_cache[dst + j] = __float2bfloat16(a);
          rope.v_cache[dst + HALF + j] = __float2bfloat16(bb);
        }
      }
      fence_proxy_async_smem();   // q in shared memory -> tcgen05 reads
      fence_proxy_async_all();    // K / V row in global memory -> this CTA's own TMA load of the last tile
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(q_ready);
        mbar_arrive(kv_ready);
      }
    } else {
      const bf16* qsrc = q + (size_t)b * ldq + (size_t)kvh * G * D;
      constexpr int CH = D / 8;  // 16-byte chunks per row
      for (int i = t; i < NQ * CH; i += 128) {
        const int g = i / CH, c = i % CH;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (g < G) v = *reinterpret_cast<const uint4*>(qsrc + g * D + c * 8);
        *reinterpret_cast<uint4*>(sQ + (c >> 3) * (NQ * 128) + g * 128 + (((c & 7) ^ (g & 7)) << 4)) = v;
      }
      for (int i = t; i < C::P_BYTES / 16; i += 128) reinterpret_cast<uint4*>(sP)[i] = make_uint4(0, 0, 0, 0);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_ready);
    }
    float m_used[G], l_run[G];
#pragma unroll
    for (int g = 0; g < G; ++g) { m_used[g] = -INFINITY; l_run[g] = 0.f; }

    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      uint32_t sv[16];
      tmem_ld_32x32b_x16(tmem_S + lane_sel + (j & 1) * NQ, sv);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[j & 1]);
      const bool valid = (t_begin + j) * BKV + t < ctx;
      float sc[G], mx[G];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        sc[g] = valid ? __uint_as_float(sv[g]) * scale_log2 : -INFINITY;
        mx[g] = sc[g];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int g = 0; g < G; ++g) mx[g] = fmaxf(mx[g], __shfl_xor_sync(0xffffffffu, mx[g], o));
      float* red = sRed + (j & 1) * 4 * kMaxG;
      if (lane == 0)
#pragma unroll
        for (int g = 0; g < G; ++g) red[qd * kMaxG + g] = mx[g];
      bar_sync(1, 128);
      bool rescale = false;
      float f[G];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float tm = fmaxf(fmaxf(red[g], red[kMaxG + g]), fmaxf(red[2 * kMaxG + g], red[3 * kMaxG + g]));
        f[g] = 1.f;
        if (tm > m_used[g] + kRescaleThreshold) {  // also the first tile (m_used = -inf); tm is finite: position 0 of the tile is valid
          f[g] = (m_used[g] == -INFINITY) ? 0.f : fast_exp2(m_used[g] - tm);
          m_used[g] = tm;
          rescale = true;
        }
      }
      float pr[G], sum[G];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        pr[g] = fast_exp2(sc[g] - m_used[g]);
        sum[g] = pr[g];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int g = 0; g < G; ++g) sum[g] += __shfl_xor_sync(0xffffffffu, sum[g], o);
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);  // P buffer free, O^T consistent
        tc_fence_after();
      }
      if (rescale && j > 0) {
        uint32_t o[16];
        tmem_ld_32x32b_x16(tmem_O + lane_sel, o);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < G; ++g) o[g] = __float_as_uint(__uint_as_float(o[g]) * f[g]);
        tmem_st_32x32b_x16(tmem_O + lane_sel, o);
        tmem_st_wait();
      }
      // P^T[g][kv = t] -> K-major swizzled B operand (row g, 128 B per 64 kv)
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int c = (t & 63) >> 3;
        *reinterpret_cast<bf16*>(sP + (t >> 6) * (NQ * 128) + g * 128 + ((c ^ (g & 7)) << 4) + (t & 7) * 2) =
            __float2bfloat16(pr[g]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      // running sum: each thread keeps only ITS warp's partial (merged across warps at the end)
#pragma unroll
      for (int g = 0; g < G; ++g) l_run[g] = l_run[g] * f[g] + sum[g];
    }
    // ---- epilogue: O^T lane d, column g; l = sum over the 4 warps' partials
    mbar_wait(pv_done, (n_tiles - 1) & 1);
    tc_fence_after();
    float* redl = sRed;  // all tiles done: reuse
    bar_sync(1, 128);
    if (lane == 0)
#pragma unroll
      for (int g = 0; g < G; ++g) redl[qd * kMaxG + g] = l_run[g];
    bar_sync(1, 128);
    uint32_t o[16];
    tmem_ld_32x32b_x16(tmem_O + lane_sel, o);
    tmem_ld_wait();
    if (out_direct != nullptr) {  // single split: this CTA saw the whole context — normalise and store, no combine pass
      if (t < D) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float lg = redl[g] + redl[kMaxG + g] + redl[2 * kMaxG + g] + redl[3 * kMaxG + g];
          out_direct[(size_t)b * ldo + (size_t)(kvh * G + g) * D + t] =
              __float2bfloat16(lg > 0.f ? __uint_as_float(o[g]) / lg : 0.f);
        }
      }
    } else {
      if (t < D) {
#pragma unroll
        for (int g = 0; g < G; ++g) ws_base[g * (D + 2) + t] = __uint_as_float(o[g]);
      }
      if (t < G) {
        ws_base[t * (D + 2) + D] = m_used[t];
        ws_base[t * (D + 2) + D + 1] = redl[t] + redl[kMaxG + t] + redl[2 * kMaxG + t] + redl[3 * kMaxG + t];
      }
    }
  }
  if (dep_done && warp >= 2) {  // the 128 softmax threads have stored this CTA's output rows: report to the next kernel
    __threadfence();
    bar_sync(1, 128);
    if (threadIdx.x == 64) dep_signal(dep_done);
  }
  if (threadIdx.x == 64) trace_ev(trace, 7);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

template <int D>
__global__ void __launch_bounds__(D / 2)
attn_decode_combine_kernel(const float* __restrict__ ws, bf16* __restrict__ out, int ldo, int Hq, int Hkv, int G,
                           int num_splits) {
  pdl_launch_dependents();
  pdl_wait();
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
cudaError_t set_attr() {
  return cudaFuncSetAttribute(attn_decode_kernel<D, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, DCfg<D>::SMEM);
}

template <int D, int G>
cudaError_t launch(cudaStream_t stream, const AttnDecodeArgs& a) {
  CUtensorMap mk, mv;
  const uint64_t blocks = (uint64_t)a.num_pages * a.Hkv;
  if (!make_tmap_3d(&mk, a.k_cache, TM_BF16, D, PAGE, blocks, (uint64_t)D * 2, (uint64_t)PAGE * D * 2, 64, PAGE, 1)) return cudaErrorInvalidValue;
  if (!make_tmap_3d(&mv, a.v_cache, TM_BF16, D, PAGE, blocks, (uint64_t)D * 2, (uint64_t)PAGE * D * 2, 64, PAGE, 1)) return cudaErrorInvalidValue;
  dim3 grid(a.num_splits, a.Hkv, a.B);
  bf16* direct = a.num_splits == 1 ? a.out : nullptr;  // one split per (sequence, kv head): no partials to merge
  const size_t ctas = (size_t)a.num_splits * a.Hkv * a.B;
  const int bank_tiles = (int)std::min<size_t>(a.sig.bank_bytes / (ctas * 2 * DCfg<D>::KV_BYTES), 64);
  cudaError_t e = launch_k(attn_decode_kernel<D, G>, grid, dim3(kThreads), DCfg<D>::SMEM, stream, true, mk, mv, a.q, a.ldq,
                           a.page_table, a.max_pages, a.ctx_lens, a.workspace, a.Hkv, a.num_splits,
                           a.scale * 1.4426950408889634f, direct, a.ldo, a.sig.wait, a.sig.wait_count, a.sig.done, bank_tiles,
                           a.k_cache, a.v_cache, a.sig.trace, a.sig.dep.wait, a.sig.dep.wait_count, a.sig.dep.done,
                           a.num_splits == 1 ? a.rope : DecodeRope{});
  if (e != cudaSuccess || direct) return e;
  return launch_k(attn_decode_combine_kernel<D>, dim3(a.Hq, a.B), dim3(D / 2), 0, stream, true,
                  (const float*)a.workspace, a.out, a.ldo, a.Hq, a.Hkv, G, a.num_splits);
}

}  // namespace

cudaError_t attn_decode_init() {
  cudaError_t e;
  if ((e = set_attr<128, 1>()) != cudaSuccess) return e;
  if ((e = set_attr<128, 2>()) != cudaSuccess) return e;
  if ((e = set_attr<128, 3>()) != cudaSuccess) return e;
  if ((e = set_attr<128, 4>()) != cudaSuccess) return e;
  if ((e = set_attr<128, 6>()) != cudaSuccess) return e;
  if ((e = set_attr<128, 8>()) != cudaSuccess) return e;
  if ((e = set_attr<64, 1>()) != cudaSuccess) return e;
  if ((e = set_attr<64, 2>()) != cudaSuccess) return e;
  if ((e = set_attr<64, 3>()) != cudaSuccess) return e;
  if ((e = set_attr<64, 4>()) != cudaSuccess) return e;
  if ((e = set_attr<64, 6>()) != cudaSuccess) return e;
  return set_attr<64, 8>();
}

size_t attn_decode_workspace_floats(int B, int Hq, int D, int num_splits) {
  return (size_t)B * Hq * num_splits * (D + 2);
}

cudaError_t attn_decode(cudaStream_t stream, const AttnDecodeArgs& a) {
  if (a.B <= 0) return cudaSuccess;
  if (a.page_size != PAGE || a.Hq % a.Hkv || a.num_splits < 1 || a.num_pages <= 0) return cudaErrorInvalidValue;
  const int G = a.Hq / a.Hkv;
  if (a.D == 128) {
    switch (G) {
      case 1: return launch<128, 1>(stream, a);
      case 2: return launch<128, 2>(stream, a);
      case 3: return launch<128, 3>(stream, a);
      case 4: return launch<128, 4>(stream, a);
      case 6: return launch<128, 6>(stream, a);
      case 8: return launch<128, 8>(stream, a);
      default: return cudaErrorInvalidValue;
    }
  }
  if (a.D == 64) {
    switch (G) {
      case 1: return launch<64, 1>(stream, a);
      case 2: return launch<64, 2>(stream, a);
      case 3: return launch<64, 3>(stream, a);
      case 4: return launch<64, 4>(stream, a);
      case 6: return launch<64, 6>(stream, a);
      case 8: return launch<64, 8>(stream, a);
      default: return cudaErrorInvalidValue;
    }
  }
  return cudaErrorInvalidValue;
}

}  // namespace hb
