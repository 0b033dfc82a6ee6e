// Decode-step GEMM (K3/K7/K8/K9/K10 at M = batch <= 256): HBM-bound weight streaming.
//
// swap-AB: the WEIGHT tile [128 rows x 64 k] is the UMMA "A" operand (all 128 TMEM lanes carry
// useful output features), the activations X[M<=256, K] are the "B" operand (UMMA N = M padded to 32),
// so D[n, m] = sum_k W[n,k] X[m,k] lands in TMEM as lane = output feature, column = batch row.
//
// stream-K: the (n-tile, k-block) grid is cut into gridDim.x equal contiguous ranges — every SM streams
// the same number of weight bytes through a deep TMA ring (>=160 KB of weights in flight per SM).
// A tile cut across CTAs is written as fp32 partial "slabs" ws[seg][m][n] (plain stores, fixed slab
// order -> bit-deterministic, no atomics); the consumer row kernel (decode_rowops.cu) sums the slabs
// and fuses the epilogue (residual+RMSNorm, RoPE+KV write, SwiGLU, logits).
#include <stdio.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "kernels.h"
#include "launch.h"
#include "ptx.cuh"
#include "tma_host.h"

namespace hb {
namespace {

constexpr int BLOCK_N = 128;  // weight rows per tile (UMMA M)
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int kThreads = 192;
// batch <= 64: ~110 KB so two CTAs share an SM (the NEXT kernel's CTA prefetches its weights while this one drains);
// larger batches need the whole SM for a deep enough ring (the activation tile alone is 16-32 KB per stage)
constexpr int smem_budget(int mpad) { return (mpad <= 64 ? 110 : 227) * 1024; }

template <int MPAD>
struct SCfg {
  static constexpr int W_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int X_BYTES = MPAD * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = W_BYTES + X_BYTES;
  static constexpr int FIXED = 1024 + 512 + 1024 + 4096;  // alignment + barriers + finisher scratch: rstd[256], red[4][256]
  static constexpr int STAGES_RAW = (smem_budget(MPAD) - FIXED) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 12 ? 12 : STAGES_RAW;
  static constexpr int SMEM = STAGES * STAGE_BYTES + FIXED;
  static constexpr int TMEM_COLS = (2 * MPAD <= 64) ? 64 : (2 * MPAD <= 128) ? 128 : (2 * MPAD <= 256) ? 256 : 512;
};

__host__ __device__ inline long long range_start(long long c, long long U, long long G) { return c * U / G; }
// CTA whose contiguous range contains linear unit u
__host__ __device__ inline int owner_of(long long u, long long U, int G) {
  int c = (int)(u * G / U);
  if (c >= G) c = G - 1;
  while (c + 1 < G && range_start(c + 1, U, G) <= u) ++c;
  while (c > 0 && range_start(c, U, G) > u) --c;
  return c;
}

__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16(x)); }

// Slab sums of one column for rows m0 + i*stride (i < R): all R x ns loads are independent of each other and of any store,
// so they are in flight together (one L2 round trip per finisher instead of one per row).  Fixed slab order 0..ns-1.
template <int R>
__device__ __forceinline__ void slab_sums(const float* __restrict__ ws, const uint8_t* __restrict__ segs, int M, int N, int n,
                                          int m0, int stride, float (&v)[R]) {
#pragma unroll
  for (int i = 0; i < R; ++i) v[i] = 0.f;
  if (n >= N) return;
  const int ns = segs[n >> 7];
  for (int s = 0; s < ns; ++s) {
    const float* base = ws + (size_t)s * M * N + n;
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int m = m0 + i * stride;
      if (m < M) v[i] += __ldcg(base + (size_t)m * N);
    }
  }
}

// Runs on the 128 epilogue threads (et = 0..127) of the CTA that delivered the last slab of `group`; named barrier 2.
__device__ __noinline__ void finish_tile(const SkinnyEpi& epi, const uint8_t* __restrict__ segs, const float* __restrict__ ws,
                                         int M, int N, int group, int et, float* s_rstd, float* s_red) {
  constexpr int R = 16;
  auto rstd = [&](int m) { return epi.ss_in ? s_rstd[m] : 1.0f; };
  if (epi.mode == SK_F32) {
    const int n = group * BLOCK_N + et;
    for (int m0 = 0; m0 < M; m0 += R) {
      float v[R];
      slab_sums<R>(ws, segs, M, N, n, m0, 1, v);
      if (n < N) {
#pragma unroll
        for (int i = 0; i < R; ++i)
          if (m0 + i < M) epi.out_f32[(size_t)(m0 + i) * epi.ldo + n] = v[i] * rstd(m0 + i);
      }
    }
  } else if (epi.mode == SK_RESID_NORM) {
    const int n = group * BLOCK_N + et;  // N % 128 == 0 (checked on the host)
    const float g = __bfloat162float(epi.gain[n]);
    const int w = et >> 5, l = et & 31;
    for (int m0 = 0; m0 < M; m0 += R) {
      float v[R], xo[R];
      slab_sums<R>(ws, segs, M, N, n, m0, 1, v);
#pragma unroll
      for (int i = 0; i < R; ++i) xo[i] = (m0 + i < M) ? __bfloat162float(epi.x[(size_t)(m0 + i) * N + n]) : 0.f;
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int m = m0 + i;
        if (m >= M) break;  // uniform over the 128 threads
        const size_t o = (size_t)m * N + n;
        const float xr = bf16r(xo[i] + v[i]);  // the ROUNDED residual feeds the norm (as in prefill)
        epi.x[o] = __float2bfloat16(xr);
        epi.xg[o] = __float2bfloat16(xr * g);
        float sq = xr * xr;
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, sh);
        if (l == 0) s_red[w * 256 + m] = sq;
      }
    }
    bar_sync(2, 128);
    for (int m = et; m < M; m += 128)
      epi.ss_out[(size_t)group * kSkinnySsStride + m] = (s_red[m] + s_red[256 + m]) + (s_red[512 + m] + s_red[768 + m]);
  } else if (epi.mode == SK_SWIGLU) {
    const int ng = (2 * group) * BLOCK_N + et, nu = ng + BLOCK_N;
    for (int m0 = 0; m0 < M; m0 += R) {
      float gv[R], uv[R];
      slab_sums<R>(ws, segs, M, N, ng, m0, 1, gv);
      slab_sums<R>(ws, segs, M, N, nu, m0, 1, uv);
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int m = m0 + i;
        if (m < M) {
          const float r = rstd(m), a = gv[i] * r, b = uv[i] * r;
          epi.h[(size_t)m * epi.F + group * BLOCK_N + et] = __float2bfloat16(a / (1.0f + __expf(-a)) * b);
        }
      }
    }
  } else if (epi.mode == SK_QKV_ROPE) {
    // 64 rotary pairs per 128-column tile (one head at D=128, two at D=64); thread = pair x row parity
    const int D = epi.D, half = D >> 1;
    const int p = et & 63, par = et >> 6;
    const int hh = p / half, j = p % half;            // head within the tile, pair index within the head
    const int c0 = hh * D + j, c1 = c0 + half;
    const int head = group * (BLOCK_N / D) + hh;
    const int n0 = group * BLOCK_N + c0, n1 = group * BLOCK_N + c1;
    const bool rot = head < epi.Hq + epi.Hkv;
    const float freq = rot ? epi.inv_freq[j] : 0.f;
    for (int m0 = par; m0 < M; m0 += 2 * R) {
      float va[R], vb[R];
      slab_sums<R>(ws, segs, M, N, n0, m0, 2, va);
      slab_sums<R>(ws, segs, M, N, n1, m0, 2, vb);
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int m = m0 + 2 * i;
        if (m >= M) continue;
        const float r = rstd(m);
        // the projection output is rounded to bf16 first (as in the prefill epilogue); RoPE is evaluated in fp32 on top
        const float a = bf16r(va[i] * r), b = bf16r(vb[i] * r);
        const int slot = epi.slots[m];
        const int page = slot >= 0 ? slot / epi.page_size : 0, off = slot >= 0 ? slot % epi.page_size : 0;
        if (rot) {
          float sn, cs;
          sincosf((float)epi.positions[m] * freq, &sn, &cs);
          const bf16 lo = __float2bfloat16(rope_lo(a, b, cs, sn)), hi = __float2bfloat16(rope_hi(a, b, cs, sn));
          if (head < epi.Hq) {
            bf16* qrow = epi.qkv_out + (size_t)m * N + (size_t)head * D;
            qrow[j] = lo;
            qrow[half + j] = hi;
          } else if (slot >= 0) {
            const size_t dst = (((size_t)page * epi.Hkv + (head - epi.Hq)) * epi.page_size + off) * D;
            epi.k_cache[dst + j] = lo;
            epi.k_cache[dst + half + j] = hi;
          }
        } else if (slot >= 0) {
          const size_t dst = (((size_t)page * epi.Hkv + (head - epi.Hq - epi.Hkv)) * epi.page_size + off) * D;
          epi.v_cache[dst + j] = __float2bfloat16(a);
          epi.v_cache[dst + half + j] = __float2bfloat16(b);
        }
      }
    }
  }
}

template <int MPAD>
__global__ void __launch_bounds__(kThreads, (MPAD <= 64 ? 2 : 1))
gemm_skinny_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x,
                   float* __restrict__ ws, int M, int N, int K, const int* sig_wait, int sig_wait_count, int* sig_done,
                   int bank_units, const bf16* __restrict__ Wp, int ldw, unsigned long long* trace,
                   const __grid_constant__ SkinnyEpi epi, const uint8_t* __restrict__ segs, const int* dep_wait, int dep_count,
                   int* dep_done) {
  using C = SCfg<MPAD>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w = smem;
  uint8_t* smem_x = smem + STAGES * C::W_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  int* s_last = reinterpret_cast<int*>(bars + 2 * STAGES + 5);
  float* s_rstd = reinterpret_cast<float*>(smem + STAGES * C::STAGE_BYTES + 512);  // [256]
  float* s_red = s_rstd + 256;                                                      // [4][256]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = (K + BLOCK_K - 1) / BLOCK_K;
  const int n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  const long long U = (long long)n_tiles * KB;
  const int G = gridDim.x;
  const long long u0 = range_start(blockIdx.x, U, G), u1 = range_start(blockIdx.x + 1, U, G);

  pdl_launch_dependents();
  if (blockIdx.x != 0) trace = nullptr;
  if (threadIdx.x == 0) {
    trace_ev(trace, 0);
    tma_prefetch_desc(&map_w);
    tma_prefetch_desc(&map_x);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);  // warp-uniform for the compiler too

  // Single-thread roles: whole warp converged through loops and waits, TMA / tcgen05 instructions predicated on an
  // elect.sync leader (under `if (lane == 0)` ptxas wraps each of them in a waterfall loop; see gemm.cu).
  if (warp == 0) {
    const bool elected = elect_one_sync();
    {
      // Weights are static: fill the whole ring with W tiles BEFORE waiting for the producer of X
      // (programmatic dependent launch) - the HBM stream never stops between back-to-back projections.
      const long long pre_end = min(u1, u0 + STAGES);
      for (long long u = u0; u < pre_end; ++u) {
        const int s = (int)(u - u0);
        if (elected) {
          mbar_arrive_expect_tx(&full_bar[s], C::STAGE_BYTES);
          tma_load_2d(smem_w + s * C::W_BYTES, &map_w, &full_bar[s], (int)(u % KB) * BLOCK_K, (int)(u / KB) * BLOCK_N, kEvictFirst);
        }
      }
      if (pre_end >= u1 && elected) sig_add(sig_done);  // the whole range fitted into the ring
      if (elected) trace_ev(trace, 1);
      if (dep_wait) {  // the activations are ready when every CTA of the previous kernel has reported (ptx.cuh dep_*)
        if (elected) dep_wait_thread(dep_wait, dep_count);
        __syncwarp();
        if (elected) fence_proxy_async_all();  // their generic-proxy stores, read below through TMA
      } else {
        pdl_wait();
      }
      if (elected) trace_ev(trace, 3);
      for (long long u = u0; u < pre_end; ++u) {
        const int s = (int)(u - u0);
        if (elected) tma_load_2d(smem_x + s * C::X_BYTES, &map_x, &full_bar[s], (int)(u % KB) * BLOCK_K, 0, kEvictLast);
      }
      int s = 0;
      uint32_t phase = 1;  // the ring has been filled once
      if (pre_end - u0 < STAGES) { s = (int)(pre_end - u0); phase = 0; }
      for (long long u = pre_end; u < u1; ++u) {
        const int tile = (int)(u / KB), kb = (int)(u % KB);
        mbar_wait(&empty_bar[s], phase ^ 1);
        if (elected) {
          mbar_arrive_expect_tx(&full_bar[s], C::STAGE_BYTES);
          tma_load_2d(smem_w + s * C::W_BYTES, &map_w, &full_bar[s], kb * BLOCK_K, tile * BLOCK_N, kEvictFirst);
          tma_load_2d(smem_x + s * C::X_BYTES, &map_x, &full_bar[s], kb * BLOCK_K, 0, kEvictLast);
          if (u + 1 == u1) { sig_add(sig_done); trace_ev(trace, 5); }  // this CTA's HBM demand ends here
        }
        if (++s == STAGES) { s = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const bool elected = elect_one_sync();
    {
      constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_N, MPAD);
      int s = 0;
      uint32_t phase = 0;
      int it = 0;
      long long u = u0;
      while (u < u1) {
        const int tile = (int)(u / KB);
        const long long seg_end = min(u1, (long long)(tile + 1) * KB);
        const int as = it & 1;
        mbar_wait(&tmem_empty[as], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * MPAD;
        bool first = true;
        for (; u < seg_end; ++u) {
          mbar_wait(&full_bar[s], phase);
          if (u == u0 && elected) trace_ev(trace, 4);
          tc_fence_after();
          const uint32_t w_addr = smem_u32(smem_w + s * C::W_BYTES);
          const uint32_t x_addr = smem_u32(smem_x + s * C::X_BYTES);
          if (elected) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              umma_f16_ss(d_tmem, umma_desc_kmajor_sw128(w_addr + k * UMMA_K * 2),
                          umma_desc_kmajor_sw128(x_addr + k * UMMA_K * 2), idesc, (first && k == 0) ? 0u : 1u);
            }
            umma_commit(&empty_bar[s]);
          }
          first = false;
          if (++s == STAGES) { s = 0; phase ^= 1; }
        }
        if (elected) umma_commit(&tmem_full[as]);
        if (elected && u >= u1) trace_ev(trace, 6);
        ++it;
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;  // weight row within the tile == TMEM lane
    int it = 0;
    long long u = u0;
    // HBM hand-over (ptx.cuh sig_*): these four warps idle until the first accumulator is ready.  Once the previous
    // streaming kernel has issued its last load, they pull the `bank_units` weight tiles that follow the ring into L2
    // (one 128-byte row of the tile per thread), so HBM keeps working through the dependency gap and the ring loads of
    // those tiles hit L2.
    if (bank_units > 0) {
      if (lane == 0) sig_wait_ge(sig_wait, sig_wait_count);
      if (warp == 2 && lane == 0) trace_ev(trace, 2);
      __syncwarp();
      const long long pre_end = min(u1, u0 + STAGES), bank_end = min(u1, pre_end + bank_units);
      for (long long ub = pre_end; ub < bank_end; ++ub) {
        const long long wrow = (ub / KB) * BLOCK_N + row;
        if (wrow < N) prefetch_l2_line(Wp + wrow * ldw + (ub % KB) * BLOCK_K);
      }
    }
    if (dep_wait) {  // ws is an activation buffer: never written before the previous kernel has finished reading it
      if (lane == 0) dep_wait_thread(dep_wait, dep_count);
      __syncwarp();
    } else {
      pdl_wait();
    }
    if (epi.mode != SK_SLABS && epi.ss_in) {
      // RMSNorm row factors of the GEMM's input rows, from the previous finisher's per-tile partials, summed in tile
      // order (deterministic).  Computed once per CTA while the first accumulator is still being produced; 8 partials
      // are loaded at a time (a dependent load-add chain would pay one L2 round trip per tile).
      for (int m = threadIdx.x - 64; m < M; m += 128) {
        float t = 0.f;
        for (int t0 = 0; t0 < epi.ss_tiles; t0 += 8) {
          float part[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            part[i] = (t0 + i < epi.ss_tiles) ? __ldcg(epi.ss_in + (size_t)(t0 + i) * kSkinnySsStride + m) : 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) t += part[i];
        }
        s_rstd[m] = rsqrtf(t / (float)epi.norm_h + epi.eps);
      }
      bar_sync(2, 128);
    }
    while (u < u1) {
      const int tile = (int)(u / KB);
      const long long seg_end = min(u1, (long long)(tile + 1) * KB);
      const int seg = blockIdx.x - owner_of((long long)tile * KB, U, G);  // slab index: fixed by the partition
      const int as = it & 1;
      mbar_wait(&tmem_full[as], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t t_tile = tmem_base + as * MPAD + (static_cast<uint32_t>(q * 32) << 16);
      const int n = tile * BLOCK_N + row;
      float* dst = ws + (size_t)seg * M * N + n;
#pragma unroll 1
      for (int c = 0; c < MPAD / 32; ++c) {
        if (c * 32 >= M) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_tile + c * 32, v);
        tmem_ld_wait();
        if (n < N) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int m = c * 32 + j;
            if (m < M) dst[(size_t)m * N] = __uint_as_float(v[j]);  // 32 lanes -> 128 contiguous bytes per m
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      if (epi.mode != SK_SLABS) {
        // arrival: the last CTA to deliver a slab of this tile (pair of tiles for SwiGLU) finishes it
        const int et = threadIdx.x - 64;  // 0..127 over the four epilogue warps
        const int group = epi.mode == SK_SWIGLU ? (tile >> 1) : tile;
        __threadfence();                  // this thread's slab stores are visible device-wide ...
        bar_sync(2, 128);                 // ... for all 128 of them before the arrival is counted
        if (et == 0) {
          const int expected = epi.mode == SK_SWIGLU ? (int)segs[2 * group] + (int)segs[2 * group + 1] : (int)segs[tile];
          const int old = atomicAdd(epi.tile_cnt + group, 1);
          const int last = old == expected - 1;
          if (last) epi.tile_cnt[group] = 0;  // nobody touches it again before the next launch
          *s_last = last;
        }
        bar_sync(2, 128);
        if (*s_last) {
          __threadfence();
          finish_tile(epi, segs, ws, M, N, group, et, s_rstd, s_red);
        }
      }
      u = seg_end;
      ++it;
    }
    if (dep_done) {  // every slab (and finisher output) of this CTA is stored: report to the next kernel
      __threadfence();
      bar_sync(2, 128);
      if (threadIdx.x == 64) dep_signal(dep_done);
    }
    if (warp == 2 && lane == 0) trace_ev(trace, 7);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

struct PlanKey {
  int dev, N, K, sms;
  bool operator<(const PlanKey& o) const { return std::tie(dev, N, K, sms) < std::tie(o.dev, o.N, o.K, o.sms); }
};
std::mutex g_plan_mu;
std::map<PlanKey, SkinnyPlan> g_plans;

template <int MPAD>
cudaError_t launch(cudaStream_t stream, const SkinnyPlan& plan, const bf16* X, int ldx, const bf16* W, int ldw,
                   float* ws, int M, int N, int K, const StreamSig* sig, const SkinnyEpi* epi) {
  using C = SCfg<MPAD>;
  const StreamSig none{};
  if (!sig) sig = &none;
  const SkinnyEpi no_epi{};
  if (!epi) epi = &no_epi;
  if (epi->mode != SK_SLABS) {
    if (!epi->tile_cnt) return cudaErrorInvalidValue;
    if (epi->mode == SK_RESID_NORM && (N % BLOCK_N)) return cudaErrorInvalidValue;
    if (epi->mode == SK_SWIGLU && (N % (2 * BLOCK_N) || epi->F * 2 != N)) return cudaErrorInvalidValue;
    if (epi->mode == SK_QKV_ROPE && ((epi->D != 64 && epi->D != 128) || (epi->Hq * epi->D) % BLOCK_N || (epi->Hkv * epi->D) % BLOCK_N ||
                                    (epi->Hq + 2 * epi->Hkv) * epi->D != N))
      return cudaErrorInvalidValue;
  }
  const int bank_units = (int)std::min<size_t>(sig->bank_bytes / ((size_t)plan.grid * C::W_BYTES), 4096);
  CUtensorMap mw, mx;
  if (!make_tmap_2d(&mw, W, TM_BF16, (uint64_t)K, (uint64_t)N, (uint64_t)ldw * 2, BLOCK_K, BLOCK_N)) return cudaErrorInvalidValue;
  if (!make_tmap_2d(&mx, X, TM_BF16, (uint64_t)K, (uint64_t)M, (uint64_t)ldx * 2, BLOCK_K, MPAD)) return cudaErrorInvalidValue;
  return launch_k(gemm_skinny_kernel<MPAD>, dim3(plan.grid), dim3(kThreads), C::SMEM, stream, true, mw, mx, ws, M, N, K,
                  sig->wait, sig->wait_count, sig->done, bank_units, W, ldw, sig->trace, *epi, plan.seg_count, sig->dep.wait,
                  sig->dep.wait_count, sig->dep.done);
}

template <int MPAD>
cudaError_t set_attr() {
  return cudaFuncSetAttribute(gemm_skinny_kernel<MPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SCfg<MPAD>::SMEM);
}

}  // namespace

cudaError_t gemm_skinny_init() {
  cudaError_t e;
  if ((e = set_attr<32>()) != cudaSuccess) return e;
  if ((e = set_attr<64>()) != cudaSuccess) return e;
  if ((e = set_attr<128>()) != cudaSuccess) return e;
  return set_attr<256>();
}

cudaError_t gemm_skinny_plan(int N, int K, SkinnyPlan* out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  sms = effective_sms(sms);  // hb_engine_cfg.sm_budget of the calling engine
  std::lock_guard<std::mutex> g(g_plan_mu);
  auto it = g_plans.find(PlanKey{dev, N, K, sms});
  if (it != g_plans.end()) {
    *out = it->second;
    return cudaSuccess;
  }
  const int KB = (K + BLOCK_K - 1) / BLOCK_K, n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  const long long U = (long long)n_tiles * KB;
  SkinnyPlan p{};
  p.n_tiles = n_tiles;
  p.grid = (int)std::min<long long>(sms, std::max<long long>(1, U / 4));
  std::vector<uint8_t> segs(n_tiles);
  int smax = 1;
  for (int t = 0; t < n_tiles; ++t) {
    const int first = owner_of((long long)t * KB, U, p.grid), last = owner_of((long long)t * KB + KB - 1, U, p.grid);
    segs[t] = (uint8_t)(last - first + 1);
    smax = std::max(smax, last - first + 1);
  }
  p.max_segs = smax;
  uint8_t* d = nullptr;
  if ((e = cudaMalloc(&d, n_tiles)) != cudaSuccess) return e;
  if ((e = cudaMemcpy(d, segs.data(), n_tiles, cudaMemcpyHostToDevice)) != cudaSuccess) return e;
  p.seg_count = d;
  g_plans[PlanKey{dev, N, K, sms}] = p;
  *out = p;
  return cudaSuccess;
}

int gemm_skinny_max_segs(int N, int K, int sms) {
  const int KB = (K + BLOCK_K - 1) / BLOCK_K, n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  const long long U = (long long)n_tiles * KB;
  const int G = (int)std::min<long long>(sms, std::max<long long>(1, U / 4));
  int smax = 1;
  for (int t = 0; t < n_tiles; ++t)
    smax = std::max(smax, owner_of((long long)t * KB + KB - 1, U, G) - owner_of((long long)t * KB, U, G) + 1);
  return smax;
}

size_t gemm_skinny_ws_floats(const SkinnyPlan& p, int M, int N) { return (size_t)p.max_segs * M * N; }

cudaError_t gemm_skinny(cudaStream_t stream, const SkinnyPlan& plan, const bf16* X, int ldx, const bf16* W, int ldw,
                        float* ws, int M, int N, int K, const StreamSig* sig, const SkinnyEpi* epi) {
  if (M <= 0 || M > 256 || (K % 8) || (ldx % 8) || (ldw % 8)) return cudaErrorInvalidValue;
  if (M <= 32) return launch<32>(stream, plan, X, ldx, W, ldw, ws, M, N, K, sig, epi);
  if (M <= 64) return launch<64>(stream, plan, X, ldx, W, ldw, ws, M, N, K, sig, epi);
  if (M <= 128) return launch<128>(stream, plan, X, ldx, W, ldw, ws, M, N, K, sig, epi);
  return launch<256>(stream, plan, X, ldx, W, ldw, ws, M, N, K, sig, epi);
}

}  // namespace hb
