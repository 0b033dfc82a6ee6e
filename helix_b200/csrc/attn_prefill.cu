// K5: varlen flash attention for prefill (causal, GQA) and for the BERT encoder
// (bidirectional), on tcgen05 tensor cores with S, P and the output accumulator O all resident in TMEM.
// See the kernel comment for the warp roles and the two-q-tile ping-pong schedule.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "kernels.h"
#include "launch.h"
#include "ptx.cuh"
#include "tma_host.h"

namespace hb {
namespace {

constexpr int BQ = 128;     // q rows per softmax warpgroup
constexpr int QPAIR = 256;  // q rows per CTA: two q tiles ping-pong on the tensor pipe
constexpr int BKV = 64;     // kv positions per tile (= one KV page)
constexpr int kThreads = 352;  // producer warp, one MMA issuer warp per q tile, two softmax warpgroups
constexpr float kRescaleThreshold = 8.0f;  // log2 units

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for two values on the FMA/ALU pipes only (no MUFU): round-to-nearest split x = n + f, f in [-0.5, 0.5], degree-3
// minimax polynomial for 2^f (relative error 7.6e-5, far below the bf16 rounding P gets), n added into the exponent
// field.  The SFU does 16 ex2/clk/SM: at 128x128 scores per kv tile that is exactly as long as the tile's two MMAs, so
// part of every row's exponentials is computed here instead (the split FlashAttention-4 uses).
__device__ __forceinline__ uint64_t exp2_poly2(uint64_t x2) {
  float x0, x1;
  f2_unpack(x2, x0, x1);
  x2 = f2_pack(fmaxf(x0, -126.f), fmaxf(x1, -126.f));  // masked scores are -inf: clamp so the exponent add cannot wrap
  const uint64_t magic = f2_pack(12582912.f, 12582912.f);  // 1.5 * 2^23: the sum's low mantissa bits hold round(x)
  const uint64_t t2 = f2_add(x2, magic);
  const uint64_t n2 = f2_add(t2, f2_pack(-12582912.f, -12582912.f));
  const uint64_t f2 = f2_fma(n2, f2_pack(-1.f, -1.f), x2);
  uint64_t p2 = f2_fma(f2_pack(0.05520550534f, 0.05520550534f), f2, f2_pack(0.24261397123f, 0.24261397123f));
  p2 = f2_fma(p2, f2, f2_pack(0.69325476885f, 0.69325476885f));
  p2 = f2_fma(p2, f2, f2_pack(0.99992769957f, 0.99992769957f));
  float t0, t1, p0, p1;
  f2_unpack(t2, t0, t1);
  f2_unpack(p2, p0, p1);
  const uint32_t r0 = __float_as_uint(p0) + (__float_as_uint(t0) << 23);
  const uint32_t r1 = __float_as_uint(p1) + (__float_as_uint(t1) << 23);
  return f2_pack(__uint_as_float(r0), __uint_as_float(r1));
}

// Optional event trace (tools/attn_test -DHB_ATTN_TRACE): CTA 0 stamps clock64() at the hand-off points of q tile 0.
#ifdef HB_ATTN_TRACE
__device__ unsigned long long g_attn_trace[8][4096];
__device__ int g_attn_trace_n[8];
#define HB_TRACE(ev)                                                          \
  do {                                                                        \
    if (blockIdx.x == 0) {                                                    \
      const int i_ = g_attn_trace_n[ev];                                      \
      if (i_ < 4096) { g_attn_trace[ev][i_] = clock64(); g_attn_trace_n[ev] = i_ + 1; } \
    }                                                                         \
  } while (0)
#else
#define HB_TRACE(ev) do {} while (0)
#endif

template <int D>
struct ACfg {
  static constexpr int Q_BYTES = BQ * D * 2;    // one q tile
  static constexpr int KV_BYTES = BKV * D * 2;
  static constexpr int STAGES = D == 128 ? 3 : 4;  // K ring and V ring depth (16 KB tiles at D = 128)
  // D = 64 (encoder; short sequences, an item's output write-back is a large part of its time) stages O in shared
  // memory and writes it with TMA: one 32-row x 128 B swizzled box per softmax warp.  At D = 128 the K/V rings leave no room.
  static constexpr bool STAGE_O = true;
  static constexpr int O_STAGE_BYTES = 32 * D * 2;  // per softmax warp
  static constexpr int BAR_BYTES = 1024;            // barriers + TMEM pointer, padded so the staging area stays 1024-aligned
  static constexpr int SMEM = 2 * Q_BYTES + 2 * STAGES * KV_BYTES + 1024 + BAR_BYTES + (STAGE_O ? 8 * O_STAGE_BYTES : 0);
  static constexpr int SUB = D / 64;  // 64-column swizzle sub-tiles per row
};

// PERSISTENT kernel: one CTA per SM walks a static list of (256-row q pair, sequence, q-head) items, heaviest
// (latest, most kv tiles under the causal mask) first, so the TMEM allocation, barrier setup and the latency of an
// item's first loads and of its output write-back overlap with the neighbouring items' work.
//
// kv tiles are 64 positions (= one KV page) and every q tile owns TWO S buffers in TMEM:
//   TMEM columns: [S0a S0b S1a S1b] 4 x 64 fp32 | O0 | O1 (D fp32 each, at 256 + t*128).  P_t(j), packed bf16x2,
//   overwrites the first 32 columns of the buffer S_t(j) was read from.
// S_t(j+1) = Q_t·K(j+1)ᵀ is therefore issued BEFORE the softmax of tile j has produced P_t(j): the softmax warps find
// their next S tile waiting when they finish one, and the tensor pipe always has the other buffer's / other q tile's
// work queued.  (With one S buffer per q tile — the previous version — S(j+1) had to wait for PV(j), i.e. for the
// softmax: ncu showed the softmax warps waiting for S 56 % of the time and the tensor pipe 41 % busy.)
//   warp 0 lane 0 : TMA producer (Q pair of the next item as soon as the last S MMA of the current one retired;
//                   K_j / V_j through KST-stage rings that keep running across items)
//   warp 1 lane 0 : MMA issuer, program order  S_t(0) S_t(1) | PV0(j) S0(j+2) PV1(j) S1(j+2) ...  P is consumed straight
//                   from TMEM (tcgen05.mma A-from-TMEM), V straight from its [kv][d] layout (MN-major B).
//   warps 2..5 / 6..9 : softmax warpgroup of q tile 0 / 1, one thread per q row; lazy O rescale; O/l -> global.
//
// PAGED = false: K/V rows come from the same packed [T, ...] activation as Q (whole-prompt prefill, encoder), kv_len == q_len.
// PAGED = true : the q rows are the LAST q_len positions of a kv_len-long sequence whose K/V (including the chunk's own,
//                written by the RoPE/KV-write kernel just before) live in the paged pool (chunked prefill, prefix-cache
//                hits): tile j is page_table[b][j].
template <int D, bool PAGED, int POLY>
__global__ void __launch_bounds__(kThreads, 1)
attn_prefill_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_o,
                    bf16* __restrict__ out, int ldo,
                    const int32_t* __restrict__ cu_seqlens, int B, int Hq, int group, int causal, float scale_log2,
                    int max_q_pairs, const int32_t* __restrict__ kv_lens, const int32_t* __restrict__ page_table,
                    int max_pages, int Hkv) {
  using C = ACfg<D>;
  constexpr int KST = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                          // [2][Q_BYTES]
  uint8_t* sK = sQ + 2 * C::Q_BYTES;           // [KST][KV_BYTES]
  uint8_t* sV = sK + KST * C::KV_BYTES;        // [KST][KV_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + KST * C::KV_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* s_full = bars + 2;    // [q tile][S buffer]
  uint64_t* p_full = bars + 6;    // [q tile][S buffer]
  uint64_t* pv_done = bars + 10;  // [q tile][tile parity]: two per q tile, so that a waiter can never be a whole
                                  // phase behind (PV(x-2) is known complete whenever PV(x) is waited for, PV(x-1) is not)
  uint64_t* o_free = bars + 14;   // [2]
  uint64_t* k_full = bars + 16;   // [KST]
  uint64_t* k_empty = k_full + KST;
  uint64_t* v_full = k_empty + KST;
  uint64_t* v_empty = v_full + KST;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(v_empty + KST);
  uint8_t* sO = reinterpret_cast<uint8_t*>(bars) + C::BAR_BYTES;  // [8 softmax warps][32 rows][D] (STAGE_O only)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_items = max_q_pairs * B * Hq;

  struct Item {
    int b, h, kvh, q0, seq0, len, kv_len, q_off, n0, n1, n;
    bool valid, act1;
  };
  auto get_item = [&](int idx) {
    Item it;
    const int per_qp = B * Hq;
    const int qp = max_q_pairs - 1 - idx / per_qp;  // heaviest (latest) q rows first, one "round" per q pair index
    const int rem = idx % per_qp;
    it.b = rem / Hq;
    it.h = rem % Hq;
    it.kvh = it.h / group;
    it.seq0 = cu_seqlens[it.b];
    it.len = cu_seqlens[it.b + 1] - it.seq0;  // q rows of this sequence in the step
    it.kv_len = PAGED ? kv_lens[it.b] : it.len;
    it.q_off = it.kv_len - it.len;  // position of the first q row (0 unless this is a later chunk of a long prompt)
    it.q0 = qp * QPAIR;
    it.valid = it.q0 < it.len;
    const int kv_tiles = (it.kv_len + BKV - 1) / BKV;
    it.act1 = it.q0 + BQ < it.len;
    // causal: q tile t needs kv tiles up to the one holding its last row's own position
    it.n0 = causal ? min(kv_tiles, (it.q_off + it.q0 + BQ - 1) / BKV + 1) : kv_tiles;
    it.n1 = it.act1 ? (causal ? min(kv_tiles, (it.q_off + it.q0 + 2 * BQ - 1) / BKV + 1) : kv_tiles) : 0;
    it.n = max(it.n0, it.n1);
    return it;
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
    if constexpr (C::STAGE_O) tma_prefetch_desc(&map_o);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 2);  // both issuers
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&pv_done[i], 1);
    }
    for (int i = 0; i < 2; ++i) mbar_init(&o_free[i], 4);
    for (int i = 0; i < KST; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 2);  // released by both issuers
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 2);
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
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);  // warp-uniform for the compiler too

  // The two single-thread roles run their loops and barrier waits with the WHOLE warp converged and only predicate the
  // TMA / tcgen05 instructions on an elect.sync leader: ptxas then keeps descriptors and addresses in uniform registers.
  // Under `if (lane == 0)` it wraps every UTCHMMA / UTMALDG in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop —
  // ~140 cycles per MMA issued, more than a 128x64x16 MMA takes to execute (ncu: tensor pipe 33-41 % busy, issuer never
  // waiting for its inputs).
  if (warp == 0) {
    const bool leader = elect_one_sync();
    {
      int kt = 0, qi = 0;  // running kv-tile / item counters: barrier phases continue across items
      for (int idx = blockIdx.x; idx < num_items; idx += gridDim.x) {
        const Item it = get_item(idx);
        if (!it.valid) continue;
        mbar_wait(q_empty, (qi & 1) ^ 1);  // every S MMA of the previous item has retired: the Q pair buffer is free
        if (leader) {
          mbar_arrive_expect_tx(q_full, (it.act1 ? 2 : 1) * C::Q_BYTES);
          for (int t = 0; t < (it.act1 ? 2 : 1); ++t)
#pragma unroll
            for (int c = 0; c < C::SUB; ++c)
              tma_load_2d(sQ + t * C::Q_BYTES + c * (BQ * 128), &map_q, q_full, it.h * D + c * 64,
                          it.seq0 + it.q0 + t * BQ, kEvictFirst);
        }
        ++qi;
        for (int j = 0; j < it.n; ++j, ++kt) {
          const int s = kt % KST;
          const uint32_t ph = (kt / KST) & 1;
          int blk = 0;  // PAGED: (page, kv head) block of the tile's page
          if constexpr (PAGED) blk = page_table[(size_t)it.b * max_pages + j] * Hkv + it.kvh;
          mbar_wait(&k_empty[s], ph ^ 1);
          if (leader) {
            mbar_arrive_expect_tx(&k_full[s], C::KV_BYTES);
#pragma unroll
            for (int c = 0; c < C::SUB; ++c) {
              uint8_t* dst = sK + s * C::KV_BYTES + c * (BKV * 128);
              if constexpr (PAGED)
                tma_load_3d(dst, &map_k, &k_full[s], c * 64, 0, blk, kEvictLast);
              else
                tma_load_2d(dst, &map_k, &k_full[s], it.kvh * D + c * 64, it.seq0 + j * BKV, kEvictLast);
            }
          }
          mbar_wait(&v_empty[s], ph ^ 1);
          if (leader) {
            mbar_arrive_expect_tx(&v_full[s], C::KV_BYTES);
#pragma unroll
            for (int c = 0; c < C::SUB; ++c) {
              uint8_t* dst = sV + s * C::KV_BYTES + c * (BKV * 128);
              if constexpr (PAGED)
                tma_load_3d(dst, &map_v, &v_full[s], c * 64, 0, blk, kEvictLast);
              else
                tma_load_2d(dst, &map_v, &v_full[s], it.kvh * D + c * 64, it.seq0 + j * BKV, kEvictLast);
            }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // One MMA issuer warp PER q tile.  With a single in-order issuer for both tiles (PV0 S0 PV1 S1 ...) the event trace
    // (tools/attn_test -DHB_ATTN_TRACE) showed P0(j) sitting ~900 cycles before the issuer looked at it — it was blocked
    // on the other warpgroup's P — so S0(j+2) was issued only ~600 cycles before the softmax needed it and the second S
    // buffer bought nothing.  The two chains are independent except for the K/V ring slots, which are released by both
    // (empty barriers count 2; a tile one issuer does not need is released by a plain arrive once it has landed).
    // The issuer's own instruction stream is on the critical path (one warp, ~4 cycles per dependent instruction), so:
    // descriptors are a per-kernel base plus a constant (the 14-bit address field never carries), ring slots and phases
    // are running counters, and the steady state is a branch-free instantiation.
    const int t = warp - 1;
    const bool leader = elect_one_sync();
    constexpr uint32_t idesc_qk = umma_idesc_bf16(BQ, BKV, 0, 0);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, D, 0, 1);  // A = P from TMEM, B = V is MN-major
    const uint64_t qdesc = umma_desc_kmajor_sw128(smem_u32(sQ + t * C::Q_BYTES));
    const uint64_t kdesc0 = umma_desc_kmajor_sw128(smem_u32(sK));
    const uint64_t vdesc0 = umma_desc_mnmajor_sw128(smem_u32(sV), BKV * 128, 1024);
    const uint32_t tmem_s = tmem_base + t * 128, tmem_o = tmem_base + 256 + t * 128;
    uint64_t* const my_s_full = s_full + t * 2;
    uint64_t* const my_p_full = p_full + t * 2;
    uint64_t* const my_pv_done = pv_done + t * 2;
    int kslot = 0, vslot = 0;  // ring slots of the next K / V tile (this issuer walks ALL tiles of an item)
    uint32_t kphase = 0, vphase = 0;
    int ps = 0;                // absolute index (over the kernel's life) of this q tile's next PV; S buffer = index & 1
    int oi = 0, qi = 0;        // items with work for this q tile (phase of o_free) / items seen (phase of q_full)
    // S(tile x) = Q_t · Kᵀ (K in ring slot kslot) into buffer x & 1
    auto mma_s = [&](int x) {
      const uint32_t d_tmem = tmem_s + (x & 1) * BKV;
      const uint64_t kd = kdesc0 + (uint64_t)(kslot * (C::KV_BYTES >> 4));
#pragma unroll
      for (int k = 0; k < D / 16; ++k)
        umma_f16_ss(d_tmem, qdesc + (((k >> 2) * (BQ * 128) + (k & 3) * 32) >> 4),
                    kd + (((k >> 2) * (BKV * 128) + (k & 3) * 32) >> 4), idesc_qk, k != 0 ? 1u : 0u);
      umma_commit(&my_s_full[x & 1]);
      umma_commit(&k_empty[kslot]);
    };
    // O_t (+)= P(tile x) · V (V in ring slot vslot)
    auto mma_pv = [&](int x, uint32_t acc) {
      const uint32_t p_tmem = tmem_s + (x & 1) * BKV;
      const uint64_t vd = vdesc0 + (uint64_t)(vslot * (C::KV_BYTES >> 4));
#pragma unroll
      for (int k = 0; k < BKV / 16; ++k)
        umma_f16_ts(tmem_o, p_tmem + k * 8, vd + ((k * 16 * 128) >> 4), idesc_pv, (acc | k) != 0 ? 1u : 0u);
      umma_commit(&my_pv_done[x & 1]);
      umma_commit(&v_empty[vslot]);
    };
    auto k_advance = [&]() { if (++kslot == KST) { kslot = 0; kphase ^= 1; } };
    auto v_advance = [&]() { if (++vslot == KST) { vslot = 0; vphase ^= 1; } };
    // one kv tile j of the item: PV(j) if this q tile has one, S(j+2) if it has one; slots it does not use are released
    auto step = [&](auto fast_tag, bool has_pv, bool has_k, bool has_s, bool first, bool last_s) {
      constexpr bool FAST = decltype(fast_tag)::value;
      mbar_wait(&v_full[vslot], vphase);
      if (FAST || has_k) mbar_wait(&k_full[kslot], kphase);
      if (FAST || has_pv) {
        if (!FAST && first) mbar_wait(&o_free[t], (oi & 1) ^ 1);  // the previous item's O_t has been read out
        mbar_wait(&my_p_full[ps & 1], (ps >> 1) & 1);
        // P(ps) exists, so S(ps) and — issued before it — PV(ps-2) have completed: observe that pv_done phase here (it
        // costs the issuer nothing) so that no phase of the barrier passes without a wait (compute-sanitizer synccheck);
        // the softmax warps only wait on pv_done when they rescale O and at the end of an item
        if (ps >= 2) mbar_wait(&my_pv_done[ps & 1], ((ps - 2) >> 1) & 1);
        tc_fence_after();
        if (t == 0 && leader) HB_TRACE(0);  // issuer saw P0(j)
        if (leader) {
          mma_pv(ps, FAST ? 1u : (first ? 0u : 1u));
          if (FAST || has_s) mma_s(ps);  // S(j+2) reuses the buffer PV(j) reads: same-thread MMAs retire in order
          if (!FAST && last_s) umma_commit(q_empty);  // that was this q tile's last S MMA of the item
        }
        if (t == 0 && leader) HB_TRACE(1);  // PV0(j), S0(j+2) issued
        ++ps;
      } else if (leader) {
        mbar_arrive(&v_empty[vslot]);
      }
      if (!FAST && has_k && !has_s && leader) mbar_arrive(&k_empty[kslot]);
      v_advance();
      if (FAST || has_k) k_advance();
    };
    for (int idx = blockIdx.x; idx < num_items; idx += gridDim.x) {
      const Item it = get_item(idx);
      if (!it.valid) continue;
      const int nt = t == 0 ? it.n0 : it.n1, n = it.n;
      mbar_wait(q_full, qi & 1);
      // S(0), S(1): their buffers are free — the PVs that read them were issued earlier by this same thread
      for (int j = 0; j < 2 && j < n; ++j) {
        mbar_wait(&k_full[kslot], kphase);
        tc_fence_after();
        if (leader) {
          if (j < nt) mma_s(ps + j); else mbar_arrive(&k_empty[kslot]);
        }
        k_advance();
      }
      if (leader) {
        if (nt == 0) mbar_arrive(q_empty);
        else if (nt <= 2) umma_commit(q_empty);
      }
      for (int j = 0; j < n; ++j) {
        if (j > 0 && j + 3 < nt)
          step(std::true_type{}, true, true, true, false, false);
        else
          step(std::false_type{}, j < nt, j + 2 < n, j + 2 < nt, j == 0, j + 3 == nt);
      }
      oi += nt > 0;
      ++qi;
    }
  } else {
    const int t = (warp - 3) >> 2;  // q tile of this warpgroup
    const int qd = warp & 3;
    const int r = qd * 32 + lane;  // q row in tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t tO = tmem_base + lane_sel + 256 + t * 128;
    int jt = 0;  // kv tiles this warpgroup has processed so far (S buffer and barrier phases)
    for (int idx = blockIdx.x; idx < num_items; idx += gridDim.x) {
      const Item it = get_item(idx);
      if (!it.valid) continue;
      const int nt = t == 0 ? it.n0 : it.n1;
      if (nt == 0) continue;
      const int len = it.kv_len;                  // kv positions of the sequence
      const int qt0 = it.q_off + it.q0 + t * BQ;  // absolute position of the tile's first q row
      const int qpos = qt0 + r;
      const int qrow = it.q0 + t * BQ + r;        // row within the sequence's q rows of this step
      float m_used = 0.f, l = 0.f;
      for (int j = 0; j < nt; ++j) {
        const int tile = jt + j;
        const uint32_t tS = tmem_base + lane_sel + t * 128 + (tile & 1) * BKV;
        if (t == 0 && warp == 3 && lane == 0) HB_TRACE(2);  // softmax 0 starts waiting for S0(j)
        mbar_wait(&s_full[t * 2 + (tile & 1)], (tile >> 1) & 1);
        tc_fence_after();
        if (t == 0 && warp == 3 && lane == 0) HB_TRACE(3);  // got S0(j)
        const int kv0 = j * BKV;
        const bool need_mask = (kv0 + BKV > len) || (causal && (kv0 + BKV - 1 > qt0));
        const int lim = causal ? min(len - 1, qpos) : len - 1;  // last valid kv position for this row
        // ---- the S row: back-to-back TMEM loads, ONE wait, then max / exp2 / pack entirely in registers
        uint32_t sv[BKV];
#pragma unroll
        for (int c = 0; c < BKV / 32; ++c) {
          uint32_t(&dst)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[c * 32]);
          tmem_ld_32x32b_x32(tS + c * 32, dst);
        }
        tmem_ld_wait();
        if (need_mask) {  // diagonal / ragged tiles only: knock the invalid columns out once, the hot loops stay branch-free
#pragma unroll
          for (int i = 0; i < BKV; ++i) sv[i] = (kv0 + i <= lim) ? sv[i] : 0xff800000u;  // -inf
        }
        float mx8[8];  // eight independent chains instead of one long dependent FMNMX chain
#pragma unroll
        for (int i = 0; i < 8; ++i) mx8[i] = __uint_as_float(sv[i]);
#pragma unroll
        for (int i = 8; i < BKV; ++i) mx8[i & 7] = fmaxf(mx8[i & 7], __uint_as_float(sv[i]));
        const float mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])), fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
        float m_new = (mx == -INFINITY) ? m_used : mx * scale_log2;
        if (j == 0) {
          m_used = m_new;
        } else {
          const bool need = m_new > m_used + kRescaleThreshold;
          if (__any_sync(0xffffffffu, need)) {
            mbar_wait(&pv_done[t * 2 + ((tile - 1) & 1)], ((tile - 1) >> 1) & 1);  // PV(j-1), hence every PV so far, is in O_t
            tc_fence_after();
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
        // ---- p = exp2(s*scale - m) (masked entries are -inf -> 0), row sum, P (bf16x2) written over the S buffer's head.
        //      POLY of every 8 column pairs take the FMA-pipe exp2, the rest the SFU; scale/subtract and the sum are packed
        if (t == 0 && warp == 3 && lane == 0) HB_TRACE(4);  // loaded, max, rescale decision done
        const uint64_t scale2 = f2_pack(scale_log2, scale_log2), negm2 = f2_pack(-m_used, -m_used);
        uint64_t sum2a = 0, sum2b = 0;
#pragma unroll
        for (int c = 0; c < BKV / 16; ++c) {
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint64_t x2 = f2_fma(f2_pack(__uint_as_float(sv[c * 16 + 2 * i]), __uint_as_float(sv[c * 16 + 2 * i + 1])),
                                       scale2, negm2);
            uint64_t p2;
            if ((i * POLY) % 8 < POLY) {  // POLY of the 8 pairs, evenly interleaved with the SFU ones
              p2 = exp2_poly2(x2);
            } else {
              float x0, x1;
              f2_unpack(x2, x0, x1);
              p2 = f2_pack(fast_exp2(x0), fast_exp2(x1));
            }
            if (i & 1) sum2b = f2_add(sum2b, p2); else sum2a = f2_add(sum2a, p2);
            float p0, p1;
            f2_unpack(p2, p0, p1);
            pk[i] = pack_bf16x2(p0, p1);
          }
          tmem_st_32x32b_x8(tS + c * 8, pk);
        }
        {
          float s0, s1;
          f2_unpack(f2_add(sum2a, sum2b), s0, s1);
          l += s0 + s1;
        }
        if (t == 0 && warp == 3 && lane == 0) HB_TRACE(5);  // exp loop done, P stores issued
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t * 2 + (tile & 1)]);
        if (t == 0 && warp == 3 && lane == 0) HB_TRACE(6);  // P0(j) handed over
      }
      // ---- item epilogue: O / l -> global, then hand O_t back to the MMA issuer
      mbar_wait(&pv_done[t * 2 + ((jt + nt - 1) & 1)], ((jt + nt - 1) >> 1) & 1);
      tc_fence_after();
      const float inv_l = l > 0.f ? 1.0f / l : 0.f;
      bool staged = false;
      if constexpr (C::STAGE_O) {
        // Every thread writing its own 128-byte row straight to global is 32 distinct lines per store instruction: ncu
        // showed the warps spending about as long draining those stores (LSU queue, then the next TMEM load and even a
        // stack reload stuck behind them) as waiting for the last PV.  Full 32-row chunks go through a swizzled staging
        // buffer and ONE bulk tensor store per warp; the ragged last chunk of a sequence keeps the direct path.
        staged = it.q0 + t * BQ + qd * 32 + 32 <= it.len;  // warp-uniform
        if (staged) {
          uint8_t* stage = sO + (warp - 3) * C::O_STAGE_BYTES;
          if (lane == 0) tma_store_wait_read<0>();  // this warp's previous store has finished reading the buffer
          __syncwarp();
#pragma unroll
          for (int c = 0; c < D / 32; ++c) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(tO + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 w;
              w.x = pack_bf16x2(__uint_as_float(o[8 * i + 0]) * inv_l, __uint_as_float(o[8 * i + 1]) * inv_l);
              w.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv_l, __uint_as_float(o[8 * i + 3]) * inv_l);
              w.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv_l, __uint_as_float(o[8 * i + 5]) * inv_l);
              w.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv_l, __uint_as_float(o[8 * i + 7]) * inv_l);
              // 128B swizzle of the box: 16-byte chunk index ^ (row & 7); the box base is 1024-aligned and lane == row
              *reinterpret_cast<uint4*>(stage + (c >> 1) * (32 * 128) + lane * 128 + ((((c & 1) * 4 + i) ^ (lane & 7)) << 4)) = w;
            }
          }
          tc_fence_before();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
#pragma unroll
            for (int c = 0; c < C::SUB; ++c)
              tma_store_2d(&map_o, stage + c * (32 * 128), it.h * D + c * 64, it.seq0 + it.q0 + t * BQ + qd * 32);
            tma_store_commit();
            mbar_arrive(&o_free[t]);
          }
        }
      }
      if (!staged) {
        bf16* orow = out + (size_t)(it.seq0 + qrow) * ldo + it.h * D;
#pragma unroll
        for (int c = 0; c < D / 32; ++c) {
          uint32_t o[32];
          tmem_ld_32x32b_x32(tO + c * 32, o);
          tmem_ld_wait();
          if (qrow < it.len) {
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
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_free[t]);
      }
      jt += nt;
    }
    if constexpr (C::STAGE_O) {
      if (lane == 0) tma_store_wait_all<0>();  // the bulk stores read this CTA's shared memory: finish them before exit
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int D, bool PAGED, int POLY>
cudaError_t launch_p(cudaStream_t stream, const AttnPrefillArgs& a) {
  using C = ACfg<D>;
  CUtensorMap mq, mk, mv, mo;
  if (!make_tmap_2d(&mo, a.out, TM_BF16, (uint64_t)a.Hq * D, (uint64_t)a.T, (uint64_t)a.ldo * 2, 64, 32)) return cudaErrorInvalidValue;
  if (!make_tmap_2d(&mq, a.q, TM_BF16, (uint64_t)a.Hq * D, (uint64_t)a.T, (uint64_t)a.ldq * 2, 64, BQ)) return cudaErrorInvalidValue;
  if (PAGED) {
    // pool plane = [page][kv head][64 positions][D]: one 3-D block per (page, kv head), as in the decode kernel
    const uint64_t blocks = (uint64_t)a.num_pages * a.Hkv;
    if (!make_tmap_3d(&mk, a.k_cache, TM_BF16, D, 64, blocks, (uint64_t)D * 2, (uint64_t)64 * D * 2, 64, 64, 1)) return cudaErrorInvalidValue;
    if (!make_tmap_3d(&mv, a.v_cache, TM_BF16, D, 64, blocks, (uint64_t)D * 2, (uint64_t)64 * D * 2, 64, 64, 1)) return cudaErrorInvalidValue;
  } else {
    if (!make_tmap_2d(&mk, a.k, TM_BF16, (uint64_t)a.Hkv * D, (uint64_t)a.T, (uint64_t)a.ldk * 2, 64, BKV)) return cudaErrorInvalidValue;
    if (!make_tmap_2d(&mv, a.v, TM_BF16, (uint64_t)a.Hkv * D, (uint64_t)a.T, (uint64_t)a.ldv * 2, 64, BKV)) return cudaErrorInvalidValue;
  }
  auto kern = attn_prefill_kernel<D, PAGED, POLY>;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int max_q_pairs = (a.max_seqlen + QPAIR - 1) / QPAIR;
  const long long items = (long long)max_q_pairs * a.B * a.Hq;
  const int grid = (int)std::min<long long>(items, effective_sms(num_sms));  // hb_engine_cfg.sm_budget
  const float scale_log2 = a.scale * 1.4426950408889634f;
  kern<<<grid, kThreads, C::SMEM, stream>>>(mq, mk, mv, mo, a.out, a.ldo, a.cu_seqlens, a.B, a.Hq, a.Hq / a.Hkv, a.causal,
                                            scale_log2, max_q_pairs, a.kv_lens, a.page_table, a.max_pages, a.Hkv);
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

// pairs (of every 8) whose exp2 runs on the FMA pipe; HB_ATTN_POLY overrides (tuning / A-B runs)
int poly_pairs(int D) {
  static int env = -2;
  if (env == -2) {
    const char* s = getenv("HB_ATTN_POLY");
    env = s ? atoi(s) : -1;
  }
  if (env >= 0) return env;
  (void)D;
  return 2;  // measured on B200 (two issuers, 64-wide kv tiles), 0 / 2 / 4 of 8: D=128 1107 / 1171 / 1139, D=64 476 / 501 / 489 TFLOP/s
}
template <int D, bool PAGED>
cudaError_t launch(cudaStream_t stream, const AttnPrefillArgs& a) {
  switch (poly_pairs(D)) {
    case 0: return launch_p<D, PAGED, 0>(stream, a);
    case 4: return launch_p<D, PAGED, 4>(stream, a);
    default: return launch_p<D, PAGED, 2>(stream, a);
  }
}

cudaError_t attn_prefill_init() {
  cudaError_t e;
#define HB_ATTR1(D_, P_, Y_)                                                                                         \
  if ((e = cudaFuncSetAttribute(attn_prefill_kernel<D_, P_, Y_>, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                ACfg<D_>::SMEM)) != cudaSuccess)                                                     \
    return e;
#define HB_ATTR(D_, P_) HB_ATTR1(D_, P_, 0) HB_ATTR1(D_, P_, 2) HB_ATTR1(D_, P_, 4)
  HB_ATTR(128, false) HB_ATTR(128, true) HB_ATTR(64, false) HB_ATTR(64, true)
#undef HB_ATTR
#undef HB_ATTR1
  return cudaSuccess;
}

cudaError_t attn_prefill(cudaStream_t stream, const AttnPrefillArgs& a) {
  if (a.B <= 0 || a.T <= 0) return cudaSuccess;
  if (a.Hq % a.Hkv) return cudaErrorInvalidValue;
  if ((a.ldq % 8) || (a.ldk % 8) || (a.ldv % 8) || (a.ldo % 8)) return cudaErrorInvalidValue;
  const bool paged = a.k_cache != nullptr;
  if (paged) {
    if (!a.v_cache || !a.page_table || !a.kv_lens || a.max_pages <= 0 || a.num_pages <= 0 || a.page_size != 64) return cudaErrorInvalidValue;
    if (a.D == 128) return launch<128, true>(stream, a);
    if (a.D == 64) return launch<64, true>(stream, a);
    return cudaErrorInvalidValue;
  }
  if (a.D == 128) return launch<128, false>(stream, a);
  if (a.D == 64) return launch<64, false>(stream, a);
  return cudaErrorInvalidValue;
}

#ifdef HB_ATTN_TRACE
void attn_trace_dump() {
  static unsigned long long h[8][4096];
  int n[8];
  cudaMemcpyFromSymbol(h, g_attn_trace, sizeof(h));
  cudaMemcpyFromSymbol(n, g_attn_trace_n, sizeof(n));
  for (int e = 0; e < 7; ++e) {
    printf("TRACE %d %d:", e, n[e]);
    for (int i = 0; i < n[e] && i < 400; ++i) printf(" %llu", h[e][i] - h[2][0]);
    printf("\n");
  }
}
#endif

cudaError_t attn_naive_check(cudaStream_t stream, const AttnPrefillArgs& a, float* out_f32) {
  dim3 grid((a.max_seqlen + 63) / 64, a.Hq, a.B);
  attn_naive_kernel<<<grid, 64, 0, stream>>>(a.q, a.ldq, a.k, a.ldk, a.v, a.ldv, out_f32, a.ldo, a.cu_seqlens, a.Hq,
                                             a.Hq / a.Hkv, a.D, a.causal, a.scale);
  return cudaGetLastError();
}

}  // namespace hb
