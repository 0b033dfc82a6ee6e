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
  static constexpr int FIXED = 1024 + 512;
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

template <int MPAD>
__global__ void __launch_bounds__(kThreads, (MPAD <= 64 ? 2 : 1))
gemm_skinny_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x,
                   float* __restrict__ ws, int M, int N, int K, const int* sig_wait, int sig_wait_count, int* sig_done,
                   int bank_units, const bf16* __restrict__ Wp, int ldw, unsigned long long* trace) {
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
      pdl_wait();
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
    pdl_wait();  // ws is an activation buffer: never written before the predecessors are done
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
      u = seg_end;
      ++it;
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
  int dev, N, K;
  bool operator<(const PlanKey& o) const { return std::tie(dev, N, K) < std::tie(o.dev, o.N, o.K); }
};
std::mutex g_plan_mu;
std::map<PlanKey, SkinnyPlan> g_plans;

template <int MPAD>
cudaError_t launch(cudaStream_t stream, const SkinnyPlan& plan, const bf16* X, int ldx, const bf16* W, int ldw,
                   float* ws, int M, int N, int K, const StreamSig* sig) {
  using C = SCfg<MPAD>;
  const StreamSig none{};
  if (!sig) sig = &none;
  const int bank_units = (int)std::min<size_t>(sig->bank_bytes / ((size_t)plan.grid * C::W_BYTES), 4096);
  CUtensorMap mw, mx;
  if (!make_tmap_2d(&mw, W, TM_BF16, (uint64_t)K, (uint64_t)N, (uint64_t)ldw * 2, BLOCK_K, BLOCK_N)) return cudaErrorInvalidValue;
  if (!make_tmap_2d(&mx, X, TM_BF16, (uint64_t)K, (uint64_t)M, (uint64_t)ldx * 2, BLOCK_K, MPAD)) return cudaErrorInvalidValue;
  return launch_k(gemm_skinny_kernel<MPAD>, dim3(plan.grid), dim3(kThreads), C::SMEM, stream, true, mw, mx, ws, M, N, K,
                  sig->wait, sig->wait_count, sig->done, bank_units, W, ldw, sig->trace);
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
  std::lock_guard<std::mutex> g(g_plan_mu);
  auto it = g_plans.find(PlanKey{dev, N, K});
  if (it != g_plans.end()) {
    *out = it->second;
    return cudaSuccess;
  }
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
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
  g_plans[PlanKey{dev, N, K}] = p;
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
                        float* ws, int M, int N, int K, const StreamSig* sig) {
  if (M <= 0 || M > 256 || (K % 8) || (ldx % 8) || (ldw % 8)) return cudaErrorInvalidValue;
  if (M <= 32) return launch<32>(stream, plan, X, ldx, W, ldw, ws, M, N, K, sig);
  if (M <= 64) return launch<64>(stream, plan, X, ldx, W, ldw, ws, M, N, K, sig);
  if (M <= 128) return launch<128>(stream, plan, X, ldx, W, ldw, ws, M, N, K, sig);
  return launch<256>(stream, plan, X, ldx, W, ldw, ws, M, N, K, sig);
}

}  // namespace hb
