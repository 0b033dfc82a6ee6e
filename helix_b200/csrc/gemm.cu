// K3/K7/K8/K9/K10: bf16 GEMM  C[M,N] = A[M,K] · W[N,K]^T  on the 5th-gen tensor cores.
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0 lane 0 : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx-count)
//   warp 1 lane 0 : MMA issuer     (tcgen05.mma cta_group::1, M=128 x N=BLOCK_N x K=16, fp32 accum in TMEM,
//                                   double-buffered accumulators so tile i+1's MMAs overlap tile i's epilogue)
//   warps 2..9    : epilogue       (tcgen05.ld -> fused bias / GELU / residual / SwiGLU -> swizzled smem -> TMA store);
//                                   two groups of four warps (one warp per TMEM lane quadrant) take alternate 64-column
//                                   chunks with their own staging buffer, so two warps per SM sub-partition hide each
//                                   other's TMEM/global latencies and short-K tiles (BERT) are not epilogue-bound
// M, N, K tails are handled by TMA (zero fill on load, clipping on store).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "kernels.h"
#include "launch.h"
#include "ptx.cuh"
#include "tma_host.h"

namespace hb {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle span
constexpr int UMMA_K = 16;
constexpr int kNumThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two groups of four)
constexpr int kEpiThreads = 128;   // per epilogue group
constexpr int kEpiBarrier = 1;     // named barrier id of group 0 (group 1 uses +1)
constexpr int kStagingBytes = BLOCK_M * 128;  // one 128-row x 128-byte chunk
constexpr int kSmemBudget = 227 * 1024;

// CG = CTAs cooperating on one UMMA (cta_group): with CG == 2 a CTA pair (cluster of 2 = one TPC) computes a 256 x BN
// tile; each CTA stages its own 128 rows of A and HALF of the W tile, so per-SM smem/L2 operand traffic drops by a third.
template <int BN, int CG = 1>
struct Cfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = (BN / CG) * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int FIXED = 2 * kStagingBytes + 1024 /*align slack*/ + 512 /*barriers*/;
  static constexpr int STAGES_RAW = (kSmemBudget - FIXED) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 10 ? 10 : STAGES_RAW;
  static constexpr int SMEM = STAGES * STAGE_BYTES + FIXED;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
};

// Tile order: a group of `grp` tiles along one dimension stays L2-resident (<= ~64 MB of that operand) while the other
// operand streams past it once per group; the host picks the dimension that minimises total DRAM traffic.
__device__ __forceinline__ void tile_coords(int t, int num_m, int num_n, int grp, int group_n, int& mi, int& ni) {
  const int num_a = group_n ? num_n : num_m;  // grouped dimension
  const int num_b = group_n ? num_m : num_n;  // swept dimension
  const int per_group = grp * num_b;
  const int g = t / per_group, r = t % per_group;
  const int first = g * grp;
  const int ga = min(grp, num_a - first);
  const int a = first + r % ga, b = r / ga;
  mi = group_n ? b : a;
  ni = group_n ? a : b;
}

// erf-GELU without special-function units: erf(z) = z * g(z^2) with a degree-8 polynomial g fitted on |z| <= 3.4
// (|erf error| <= 2.8e-4, |gelu error| <= 6.5e-4 — below bf16 resolution of the output), clamped to [-1, 1] beyond.
// 16 FMA-pipe instructions per element instead of libdevice erff (branchy) or rcp+ex2: the FFN1 epilogue of the
// encoder (K = 768, i.e. only 6144 tensor cycles per tile) must stay shorter than its mainloop.
// evaluated two elements per FMA-pipe issue slot (FMUL2 / FFMA2): 9.5 instead of 16 instructions per element
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
  auto c2 = [](float c) { return f2_pack(c, c); };
  const uint64_t x = f2_pack(x0, x1);
  const uint64_t z = f2_mul(x, c2(0.70710678118654752f));
  float u0, u1;
  f2_unpack(f2_mul(z, z), u0, u1);
  const uint64_t u = f2_pack(fminf(u0, 3.4f * 3.4f), fminf(u1, 3.4f * 3.4f));
  uint64_t g = f2_fma(c2(2.409683474979829e-08f), u, c2(-1.3351223060453776e-06f));
  g = f2_fma(g, u, c2(3.208911948604509e-05f));
  g = f2_fma(g, u, c2(-0.00044352986151352525f));
  g = f2_fma(g, u, c2(0.003962765447795391f));
  g = f2_fma(g, u, c2(-0.024541884660720825f));
  g = f2_fma(g, u, c2(0.11060382425785065f));
  g = f2_fma(g, u, c2(-0.3752793073654175f));
  g = f2_fma(g, u, c2(1.1283255815505981f));
  float e0, e1;
  f2_unpack(f2_mul(z, g), e0, e1);
  const uint64_t e = f2_pack(fmaxf(fminf(e0, 1.0f), -1.0f), fmaxf(fminf(e1, 1.0f), -1.0f));
  const uint64_t hx = f2_mul(x, c2(0.5f));
  f2_unpack(f2_fma(hx, e, hx), x0, x1);
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }

template <int BN, int EPI, int CG = 1>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_r,
               const bf16* __restrict__ bias, int M, int N, int K, int tile_grp, int tile_group_n,
               const __grid_constant__ RopeEpi rope) {
  using C = Cfg<BN, CG>;
  constexpr int TILE_M = BLOCK_M * CG;  // rows of one UMMA tile (per CTA: BLOCK_M)
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;
  const int tile_worker = (CG == 2) ? (blockIdx.x >> 1) : blockIdx.x;       // CTA pairs walk the tile list together
  const int tile_workers = (CG == 2) ? (gridDim.x >> 1) : gridDim.x;
  constexpr int STAGES = C::STAGES;
  constexpr bool kF32 = (EPI == EPI_F32);
  constexpr bool kSwiGLU = (EPI == EPI_SWIGLU);
  constexpr bool kResid = (EPI == EPI_RESID || EPI == EPI_BIAS_RESID);
  constexpr bool kBias = (EPI == EPI_BIAS || EPI == EPI_BIAS_GELU || EPI == EPI_BIAS_RESID);
  constexpr bool kRope = (EPI == EPI_ROPE);
  constexpr int OUT_BN = kSwiGLU ? BN / 2 : BN;          // output columns per tile
  constexpr int CHUNK_COLS = kF32 ? 32 : 64;             // output columns per 128-byte staging row
  constexpr int NCHUNK = (OUT_BN + CHUNK_COLS - 1) / CHUNK_COLS;
  static_assert(!kSwiGLU || BN == 256, "SwiGLU epilogue expects [128 gate | 128 up] tiles");
  static_assert(OUT_BN % 64 == 0, "tile width");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * C::A_BYTES;
  uint8_t* staging = smem + STAGES * C::STAGE_BYTES;  // 2 x 16 KB, 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + 2 * kStagingBytes);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;      // [2]
  uint64_t* tmem_empty = bars + 2 * STAGES + 2; // [2]
  uint64_t* resid_bar = bars + 2 * STAGES + 4;  // [2] one per epilogue group
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 6);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (M + TILE_M - 1) / TILE_M;
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_c);
    if (kResid) tma_prefetch_desc(&map_r);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], CG);   // leader: own arrive.expect_tx (+ the peer producer's remote arrive)
      mbar_init(&empty_bar[i], 1);   // tcgen05.commit (multicast to both CTAs when CG == 2)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8 * CG);  // epilogue warps of BOTH CTAs release the leader's accumulator stage
    }
    mbar_init(&resid_bar[0], 1);
    mbar_init(&resid_bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) {
      tmem_alloc_2sm(tmem_ptr, C::TMEM_COLS);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_ptr, C::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);  // warp-uniform for the compiler too

  // The single-thread roles keep the whole warp converged through their loops and barrier waits and predicate only the
  // TMA / tcgen05 instructions on an elect.sync leader: descriptors and addresses then live in uniform registers.  Under
  // `if (lane == 0)` ptxas wraps every UTCHMMA / UTMALDG in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop.
  if (warp == 0) {
    // ===================== TMA producer =====================
    const bool elected = elect_one_sync();
    // L2 policy follows the tile order: the grouped operand's slice is what the rasterisation keeps resident (evict-last),
    // the swept operand streams past it (evict-first).  Bit 1 of tile_group_n switches the hints on.
    const bool hints = (tile_group_n & 2) != 0;
    const uint64_t hint_a = !hints ? kEvictNormal : ((tile_group_n & 1) ? kEvictFirst : kEvictLast);
    const uint64_t hint_b = !hints ? kEvictNormal : ((tile_group_n & 1) ? kEvictLast : kEvictFirst);
    {
      int s = 0;
      uint32_t phase = 0;
      for (int t = tile_worker; t < num_tiles; t += tile_workers) {
        int mi, ni;
        tile_coords(t, num_m, num_n, tile_grp, tile_group_n & 1, mi, ni);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[s], phase ^ 1);
          if (elected) {
            if constexpr (CG == 2) {
              // both CTAs' bytes complete on the LEADER's barrier (its MMA thread is the only consumer)
              if (is_leader) mbar_arrive_expect_tx(&full_bar[s], 2 * C::STAGE_BYTES);
              else mbar_arrive_cluster(&full_bar[s], 0);
              tma_load_2d_2sm(smem_a + s * C::A_BYTES, &map_a, &full_bar[s], kb * BLOCK_K,
                              mi * TILE_M + (int)cta_rank * BLOCK_M, hint_a);
              tma_load_2d_2sm(smem_b + s * C::B_BYTES, &map_b, &full_bar[s], kb * BLOCK_K,
                              ni * BN + (int)cta_rank * (BN / 2), hint_b);
            } else {
              mbar_arrive_expect_tx(&full_bar[s], C::STAGE_BYTES);
              tma_load_2d(smem_a + s * C::A_BYTES, &map_a, &full_bar[s], kb * BLOCK_K, mi * BLOCK_M, hint_a);
              tma_load_2d(smem_b + s * C::B_BYTES, &map_b, &full_bar[s], kb * BLOCK_K, ni * BN, hint_b);
            }
          }
          if (++s == STAGES) { s = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (is_leader) {
      const bool elected = elect_one_sync();
      constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, BN);
      int s = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = tile_worker; t < num_tiles; t += tile_workers, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + s * C::A_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + s * C::B_BYTES);
          if (elected) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              const uint64_t adesc = umma_desc_kmajor_sw128(a_addr + k * UMMA_K * 2);
              const uint64_t bdesc = umma_desc_kmajor_sw128(b_addr + k * UMMA_K * 2);
              if constexpr (CG == 2) umma_f16_ss_2sm(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
              else umma_f16_ss(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            // frees this smem stage (in both CTAs when paired) once the MMAs above retire
            if constexpr (CG == 2) umma_commit_2sm_mc(&empty_bar[s], 0x3); else umma_commit(&empty_bar[s]);
          }
          if (++s == STAGES) { s = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if (elected) {
          if constexpr (CG == 2) umma_commit_2sm_mc(&tmem_full[as], 0x3); else umma_commit(&tmem_full[as]);
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int q = warp & 3;               // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;        // row within the 128-row tile == TMEM lane
    const int grp = (warp - 2) >> 2;      // epilogue group: chunks c == grp (mod 2), staging buffer `grp`
    const bool leader = ((threadIdx.x - 64) & 127) == 0;
    uint8_t* const buf = staging + grp * kStagingBytes;
    uint8_t* const my_row = buf + row * 128;
    uint64_t* const my_resid_bar = &resid_bar[grp];
    uint32_t resid_phase = 0;
    int it = 0;
    for (int t = tile_worker; t < num_tiles; t += tile_workers, ++it) {
      int mi, ni;
      tile_coords(t, num_m, num_n, tile_grp, tile_group_n & 1, mi, ni);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int m0 = mi * TILE_M + (int)cta_rank * BLOCK_M;
      const int n_out0 = ni * OUT_BN;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t t_tile = tmem_base + as * BN + (static_cast<uint32_t>(q * 32) << 16);

      if constexpr (kRope) {
        // QKV projection of a decoder: the tile is BN / D whole heads.  Group g takes heads g, g + 2, ...; a thread owns
        // its token's row of the head and pairs column i with i + D/2 straight from TMEM: round to bf16 (what the row
        // kernel read back), rotate by the step's cos/sin table, store to the qkv activation, and for k / v heads to the
        // row's slot of the paged cache as well.  64-byte pieces per thread; the mainloop (K = hidden) hides them.
        const int D = rope.D, half = D >> 1;
        const int row_g = m0 + row;
        const bool row_ok = row_g < M;
        const float* cs = rope.cs + (size_t)(row_ok ? row_g : 0) * D;
        const int slot = (row_ok && rope.slots) ? rope.slots[row_g] : -1;
        const int page = slot >= 0 ? slot / rope.page_size : 0, off = slot >= 0 ? slot % rope.page_size : 0;
        bf16* const orow = rope.out + (size_t)(row_ok ? row_g : 0) * rope.ldc;
        for (int hl = grp; hl < BN / D; hl += 2) {
          const int n_head0 = n_out0 + hl * D;
          const int hidx = n_head0 / D;
          const int kind = hidx < rope.Hq ? 0 : (hidx < rope.Hq + rope.Hkv ? 1 : 2);  // q / k / v
          bf16* crow = nullptr;
          if (kind != 0 && slot >= 0) {
            const int kvh = hidx - rope.Hq - (kind == 2 ? rope.Hkv : 0);
            crow = (kind == 1 ? rope.k_cache : rope.v_cache) + (((size_t)page * rope.Hkv + kvh) * rope.page_size + off) * D;
          }
          for (int hh = 0; hh < half / 32; ++hh) {
            uint32_t v1[32], v2[32];
            tmem_ld_32x32b_x32(t_tile + hl * D + hh * 32, v1);
            tmem_ld_32x32b_x32(t_tile + hl * D + half + hh * 32, v2);
            tmem_ld_wait();
            float a[32], b[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) { a[j] = __uint_as_float(v1[j]); b[j] = __uint_as_float(v2[j]); }
            if (bias) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float ba[8], bb[8];
                unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(bias + n_head0 + hh * 32) + j), ba);
                unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(bias + n_head0 + half + hh * 32) + j), bb);
#pragma unroll
                for (int k = 0; k < 8; ++k) { a[8 * j + k] += ba[k]; b[8 * j + k] += bb[k]; }
              }
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              a[j] = __bfloat162float(__float2bfloat16(a[j]));
              b[j] = __bfloat162float(__float2bfloat16(b[j]));
            }
            if (kind != 2) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 c4 = __ldg(reinterpret_cast<const float4*>(cs + hh * 32) + j);
                const float4 s4 = __ldg(reinterpret_cast<const float4*>(cs + half + hh * 32) + j);
                const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float x = a[4 * j + k], y = b[4 * j + k];
                  a[4 * j + k] = rope_lo(x, y, cc[k], ss[k]);
                  b[4 * j + k] = rope_hi(x, y, cc[k], ss[k]);
                }
              }
            }
            if (row_ok) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 lo, hi;
                lo.x = pack_bf16x2(a[8 * j + 0], a[8 * j + 1]); lo.y = pack_bf16x2(a[8 * j + 2], a[8 * j + 3]);
                lo.z = pack_bf16x2(a[8 * j + 4], a[8 * j + 5]); lo.w = pack_bf16x2(a[8 * j + 6], a[8 * j + 7]);
                hi.x = pack_bf16x2(b[8 * j + 0], b[8 * j + 1]); hi.y = pack_bf16x2(b[8 * j + 2], b[8 * j + 3]);
                hi.z = pack_bf16x2(b[8 * j + 4], b[8 * j + 5]); hi.w = pack_bf16x2(b[8 * j + 6], b[8 * j + 7]);
                reinterpret_cast<uint4*>(orow + n_head0 + hh * 32)[j] = lo;
                reinterpret_cast<uint4*>(orow + n_head0 + half + hh * 32)[j] = hi;
                if (crow) {
                  reinterpret_cast<uint4*>(crow + hh * 32)[j] = lo;
                  reinterpret_cast<uint4*>(crow + half + hh * 32)[j] = hi;
                }
              }
            }
          }
        }
        // this warp's TMEM reads of the accumulator stage are complete -> hand it back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (CG == 2) mbar_arrive_cluster(&tmem_empty[as], 0); else mbar_arrive(&tmem_empty[as]); }
        continue;
      }
      int last_c = -1;
      for (int c = grp; c < NCHUNK; c += 2) last_c = c;
      if (last_c < 0) {  // this group has no chunk in such a narrow tile: release the accumulator right away
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (CG == 2) mbar_arrive_cluster(&tmem_empty[as], 0); else mbar_arrive(&tmem_empty[as]); }
      }
#pragma unroll 1
      for (int c = grp; c < NCHUNK; c += 2) {
        if (leader) {
          tma_store_wait_read<0>();  // this group's previous store is done reading `buf`
          if (kResid) {
            mbar_arrive_expect_tx(my_resid_bar, kStagingBytes);
            tma_load_2d(buf, &map_r, my_resid_bar, n_out0 + c * CHUNK_COLS, m0, kEvictFirst);
          }
        }
        bar_sync(kEpiBarrier + grp, kEpiThreads);

        if constexpr (kF32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_tile + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 o = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            *reinterpret_cast<uint4*>(my_row + ((j ^ (row & 7)) << 4)) = o;
          }
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h) {  // two 32-column halves of the 64-column chunk
            uint32_t v[32];
            float f[32];
            const int col = c * 64 + h * 32;  // output column within tile
            tmem_ld_32x32b_x32(t_tile + col, v);
            if constexpr (kSwiGLU) {
              uint32_t u[32];
              tmem_ld_32x32b_x32(t_tile + 128 + col, u);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = silu(__uint_as_float(v[j])) * __uint_as_float(u[j]);
            } else {
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            }
            if constexpr (kBias) {
              const int nb = n_out0 + col;
              if (nb + 32 <= N) {  // warp-uniform: four 16-byte read-only loads instead of 32 scalar ones
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint4 bv = __ldg(reinterpret_cast<const uint4*>(bias + nb) + j);
                  const float2 b0 = unpack_bf16x2(bv.x), b1 = unpack_bf16x2(bv.y), b2 = unpack_bf16x2(bv.z),
                               b3 = unpack_bf16x2(bv.w);
                  f[8 * j + 0] += b0.x; f[8 * j + 1] += b0.y; f[8 * j + 2] += b1.x; f[8 * j + 3] += b1.y;
                  f[8 * j + 4] += b2.x; f[8 * j + 5] += b2.y; f[8 * j + 6] += b3.x; f[8 * j + 7] += b3.y;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const int n = nb + j;
                  f[j] += (n < N) ? __bfloat162float(bias[n]) : 0.0f;
                }
              }
            }
            if constexpr (EPI == EPI_BIAS_GELU) {
#pragma unroll
              for (int j = 0; j < 32; j += 2) gelu_erf2(f[j], f[j + 1]);
            }
            if constexpr (kResid) {
              if (h == 0) {
                mbar_wait(my_resid_bar, resid_phase);
                resid_phase ^= 1;
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 rv = *reinterpret_cast<const uint4*>(my_row + (((h * 4 + j) ^ (row & 7)) << 4));
                const float2 r0 = unpack_bf16x2(rv.x), r1 = unpack_bf16x2(rv.y), r2 = unpack_bf16x2(rv.z),
                             r3 = unpack_bf16x2(rv.w);
                f[8 * j + 0] += r0.x; f[8 * j + 1] += r0.y; f[8 * j + 2] += r1.x; f[8 * j + 3] += r1.y;
                f[8 * j + 4] += r2.x; f[8 * j + 5] += r2.y; f[8 * j + 6] += r3.x; f[8 * j + 7] += r3.y;
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o;
              o.x = pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
              o.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
              o.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
              o.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
              *reinterpret_cast<uint4*>(my_row + (((h * 4 + j) ^ (row & 7)) << 4)) = o;
            }
          }
        }
        if (c == last_c) {
          // this group's TMEM reads of the accumulator stage are complete -> hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (CG == 2) mbar_arrive_cluster(&tmem_empty[as], 0); else mbar_arrive(&tmem_empty[as]); }
        }
        fence_proxy_async_smem();
        bar_sync(kEpiBarrier + grp, kEpiThreads);
        if (leader) {
          tma_store_2d(&map_c, buf, n_out0 + c * CHUNK_COLS, m0);
          tma_store_commit();
        }
      }
    }
    if (leader) tma_store_wait_all<0>();
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_2sm(tmem_base, C::TMEM_COLS); else tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int BN, int EPI, int CG>
cudaError_t launch_cfg(cudaStream_t stream, const GemmArgs& g, int num_sms) {
  using C = Cfg<BN, CG>;
  constexpr bool kF32 = (EPI == EPI_F32);
  constexpr bool kSwiGLU = (EPI == EPI_SWIGLU);
  constexpr int TILE_M = BLOCK_M * CG;
  CUtensorMap ma, mb, mc, mr;
  const int n_out = kSwiGLU ? g.N / 2 : g.N;
  if (!make_tmap_2d(&ma, g.A, TM_BF16, (uint64_t)g.K, (uint64_t)g.M, (uint64_t)g.lda * 2, BLOCK_K, BLOCK_M)) return cudaErrorInvalidValue;
  if (!make_tmap_2d(&mb, g.W, TM_BF16, (uint64_t)g.K, (uint64_t)g.N, (uint64_t)g.ldw * 2, BLOCK_K, BN / CG)) return cudaErrorInvalidValue;
  if (kF32) {
    if (!make_tmap_2d(&mc, g.C, TM_F32, (uint64_t)n_out, (uint64_t)g.M, (uint64_t)g.ldc * 4, 32, BLOCK_M)) return cudaErrorInvalidValue;
  } else {
    if (!make_tmap_2d(&mc, g.C, TM_BF16, (uint64_t)n_out, (uint64_t)g.M, (uint64_t)g.ldc * 2, 64, BLOCK_M)) return cudaErrorInvalidValue;
  }
  if (EPI == EPI_RESID || EPI == EPI_BIAS_RESID) {
    if (!make_tmap_2d(&mr, g.R, TM_BF16, (uint64_t)n_out, (uint64_t)g.M, (uint64_t)g.ldr * 2, 64, BLOCK_M)) return cudaErrorInvalidValue;
  } else {
    mr = mc;
  }
  auto kern = gemm_tn_kernel<BN, EPI, CG>;
  static bool attr_set = false;  // per template instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int num_m = (g.M + TILE_M - 1) / TILE_M, num_n = (g.N + BN - 1) / BN;
  const int tiles = num_m * num_n;
  const int workers = std::min(tiles, num_sms / CG);
  // rasterisation: keep <= ~32 MB of one operand L2-resident, stream the other
  static const double budget = [] { const char* e = getenv("HB_GEMM_L2MB"); return (e ? atof(e) : 32.0) * 1e6; }();
  const double a_tile = (double)TILE_M * g.K * 2, w_tile = (double)BN * g.K * 2;
  const int gm = (int)std::max(1.0, std::min((double)num_m, floor(budget / a_tile)));
  const int gn = (int)std::max(1.0, std::min((double)num_n, floor(budget / w_tile)));
  const double A = a_tile * num_m, W = w_tile * num_n;
  const double t_m = A + W * ceil((double)num_m / gm), t_n = W + A * ceil((double)num_n / gn);
  int group_n = t_n < t_m ? 1 : 0;
  int grp = group_n ? gn : gm;
  static const int force_gm = [] { const char* e = getenv("HB_GEMM_GM"); return e ? atoi(e) : 0; }();  // A/B knob
  if (force_gm > 0) { group_n = 0; grp = std::min(force_gm, num_m); }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(workers * CG);
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = C::SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (CG > 1) ? 1 : 0;
  static const int l2_hints = [] { const char* e = getenv("HB_GEMM_L2HINTS"); return e ? atoi(e) : 0; }();  // A/B knob
  return cudaLaunchKernelEx(&cfg, kern, ma, mb, mc, mr, g.bias, g.M, g.N, g.K, grp, group_n | (l2_hints ? 2 : 0), g.rope);
}

bool use_pair() {  // HB_GEMM_2CTA=0 falls back to single-CTA tiles (A/B measurements)
  static const bool on = [] { const char* e = getenv("HB_GEMM_2CTA"); return !(e && !strcmp(e, "0")); }();
  return on;
}

template <int EPI>
cudaError_t launch_epi(cudaStream_t stream, const GemmArgs& g, int num_sms) {
  int bn = g.block_n;
  if (bn == 0) {
    // Fill the machine: prefer 256-wide tiles when there are at least ~2 waves of them.
    const int num_m = (g.M + BLOCK_M - 1) / BLOCK_M;
    bn = 256;
    while (bn > 64 && (long)num_m * ((g.N + bn - 1) / bn) < 2L * num_sms) bn >>= 1;
  }
  switch (bn) {
    case 256:
      // CTA pairs (256 x 256 tiles) once the problem has at least a wave of them
      if (use_pair() && (long)((g.M + 255) / 256) * ((g.N + 255) / 256) >= num_sms / 2) return launch_cfg<256, EPI, 2>(stream, g, num_sms);
      return launch_cfg<256, EPI, 1>(stream, g, num_sms);
    case 128:
      if constexpr (EPI != EPI_SWIGLU && EPI != EPI_ROPE) return launch_cfg<128, EPI, 1>(stream, g, num_sms);
      return cudaErrorInvalidValue;
    case 64:
      if constexpr (EPI != EPI_SWIGLU && EPI != EPI_ROPE) return launch_cfg<64, EPI, 1>(stream, g, num_sms);
      return cudaErrorInvalidValue;
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace

namespace {
template <int BN, int EPI, int CG>
cudaError_t set_attr() {
  return cudaFuncSetAttribute(gemm_tn_kernel<BN, EPI, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN, CG>::SMEM);
}
template <int EPI>
cudaError_t set_attr_epi() {
  cudaError_t e;
  if ((e = set_attr<256, EPI, 2>()) != cudaSuccess) return e;
  if ((e = set_attr<256, EPI, 1>()) != cudaSuccess) return e;
  if constexpr (EPI != EPI_SWIGLU && EPI != EPI_ROPE) {
    if ((e = set_attr<128, EPI, 1>()) != cudaSuccess) return e;
    return set_attr<64, EPI, 1>();
  }
  return cudaSuccess;
}
}  // namespace

cudaError_t gemm_init() {
  cudaError_t e;
  if ((e = set_attr_epi<EPI_NONE>()) != cudaSuccess) return e;
  if ((e = set_attr_epi<EPI_BIAS>()) != cudaSuccess) return e;
  if ((e = set_attr_epi<EPI_BIAS_GELU>()) != cudaSuccess) return e;
  if ((e = set_attr_epi<EPI_RESID>()) != cudaSuccess) return e;
  if ((e = set_attr_epi<EPI_BIAS_RESID>()) != cudaSuccess) return e;
  if ((e = set_attr_epi<EPI_F32>()) != cudaSuccess) return e;
  if ((e = set_attr_epi<EPI_ROPE>()) != cudaSuccess) return e;
  return set_attr_epi<EPI_SWIGLU>();
}

cudaError_t kernels_init() {
  cudaError_t e;
  if ((e = gemm_init()) != cudaSuccess) return e;
  if ((e = attn_prefill_init()) != cudaSuccess) return e;
  if ((e = gemm_skinny_init()) != cudaSuccess) return e;
  return attn_decode_init();
}

cudaError_t gemm_bf16_tn(cudaStream_t stream, const GemmArgs& g) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return cudaErrorInvalidValue;
  if ((g.K % 8) || (g.lda % 8) || (g.ldw % 8)) return cudaErrorInvalidValue;  // 16-byte TMA strides
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int sms = effective_sms(num_sms);  // hb_engine_cfg.sm_budget of the calling engine
  switch (g.epi) {
    case EPI_NONE: return launch_epi<EPI_NONE>(stream, g, sms);
    case EPI_BIAS: return launch_epi<EPI_BIAS>(stream, g, sms);
    case EPI_BIAS_GELU: return launch_epi<EPI_BIAS_GELU>(stream, g, sms);
    case EPI_RESID: return launch_epi<EPI_RESID>(stream, g, sms);
    case EPI_BIAS_RESID: return launch_epi<EPI_BIAS_RESID>(stream, g, sms);
    case EPI_SWIGLU: {
      if (g.N % 256) return cudaErrorInvalidValue;
      GemmArgs h = g;
      h.block_n = 256;
      return launch_epi<EPI_SWIGLU>(stream, h, sms);
    }
    case EPI_F32: return launch_epi<EPI_F32>(stream, g, sms);
    case EPI_ROPE: {
      const RopeEpi& r = g.rope;
      if ((g.N % 256) || !r.out || !r.cs || (r.D != 64 && r.D != 128) || g.N != (r.Hq + 2 * r.Hkv) * r.D || (r.ldc % 8)) return cudaErrorInvalidValue;
      if (r.slots && (!r.k_cache || !r.v_cache || r.page_size <= 0)) return cudaErrorInvalidValue;
      GemmArgs h = g;
      h.block_n = 256;
      return launch_epi<EPI_ROPE>(stream, h, sms);
    }
  }
  return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------ test-only checker
namespace {
__global__ void gemm_naive_kernel(const bf16* A, int lda, const bf16* W, int ldw, void* C, int ldc, int M, int N, int K,
                                  int f32) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc += __bfloat162float(A[(size_t)m * lda + k]) * __bfloat162float(W[(size_t)n * ldw + k]);
  if (f32)
    reinterpret_cast<float*>(C)[(size_t)m * ldc + n] = acc;
  else
    reinterpret_cast<bf16*>(C)[(size_t)m * ldc + n] = __float2bfloat16(acc);
}
}  // namespace

cudaError_t gemm_naive_check(cudaStream_t stream, const GemmArgs& g) {
  dim3 grid((g.N + 127) / 128, g.M);
  gemm_naive_kernel<<<grid, 128, 0, stream>>>(g.A, g.lda, g.W, g.ldw, g.C, g.ldc, g.M, g.N, g.K, g.epi == EPI_F32);
  return cudaGetLastError();
}

}  // namespace hb
