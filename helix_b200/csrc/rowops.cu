// K1/K2/K4/K10(sampling)/K11 of SURVEY.md §2.3: HBM-bound row kernels.
// 16-byte vectorised coalesced loads, warp-shuffle + one smem hop for block reductions, fp32 math.
#include <math.h>

#include "kernels.h"
#include "launch.h"
#include "ptx.cuh"

namespace hb {
namespace {

constexpr int kRowThreads = 128;
constexpr int kMaxVec = 4;  // row cached in registers up to 128*4*8 = 4096 elements

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int NT>
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();  // protect `red` reuse
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < NT / 32) ? red[l] : 0.f;
  return warp_sum(t);
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

__global__ void embed_gather_kernel(const int32_t* __restrict__ tokens, const bf16* __restrict__ table,
                                    bf16* __restrict__ x, int H) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = blockIdx.x;
  const uint4* src = reinterpret_cast<const uint4*>(table + (size_t)tokens[t] * H);
  uint4* dst = reinterpret_cast<uint4*>(x + (size_t)t * H);
  for (int i = threadIdx.x; i < H / 8; i += blockDim.x) dst[i] = __ldg(src + i);
}

__global__ void __launch_bounds__(kRowThreads)
rmsnorm_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, bf16* __restrict__ out,
               const int32_t* __restrict__ row_index, int H, float eps) {
  __shared__ float red[kRowThreads / 32];
  pdl_launch_dependents();
  pdl_wait();
  const int r = blockIdx.x;
  const size_t src_row = row_index ? (size_t)row_index[r] : (size_t)r;
  const uint4* src = reinterpret_cast<const uint4*>(x + src_row * H);
  const uint4* wv = reinterpret_cast<const uint4*>(w);
  uint4* dst = reinterpret_cast<uint4*>(out + (size_t)r * H);
  const int nvec = H / 8;
  uint4 cache[kMaxVec];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int i = threadIdx.x + j * kRowThreads;
    if (i < nvec) {
      cache[j] = src[i];
      float f[8];
      unpack8(cache[j], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) ss += f[k] * f[k];
    }
  }
  for (int i = threadIdx.x + kMaxVec * kRowThreads; i < nvec; i += kRowThreads) {
    float f[8];
    unpack8(src[i], f);
#pragma unroll
    for (int k = 0; k < 8; ++k) ss += f[k] * f[k];
  }
  ss = block_sum<kRowThreads>(ss, red);
  const float inv = rsqrtf(ss / (float)H + eps);
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int i = threadIdx.x + j * kRowThreads;
    if (i < nvec) {
      float f[8], g[8];
      unpack8(cache[j], f);
      unpack8(__ldg(wv + i), g);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = f[k] * inv * g[k];
      dst[i] = pack8(f);
    }
  }
  for (int i = threadIdx.x + kMaxVec * kRowThreads; i < nvec; i += kRowThreads) {
    float f[8], g[8];
    unpack8(src[i], f);
    unpack8(__ldg(wv + i), g);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = f[k] * inv * g[k];
    dst[i] = pack8(f);
  }
}

// LayerNorm over rows of up to kMaxVec*128*8 elements held in registers (two-pass mean / variance).
template <bool kEmbed>
__global__ void __launch_bounds__(kRowThreads)
layernorm_kernel(const bf16* __restrict__ x, const bf16* __restrict__ gamma, const bf16* __restrict__ beta,
                 bf16* __restrict__ out, int H, float eps, const int32_t* __restrict__ tokens,
                 const int32_t* __restrict__ positions, const bf16* __restrict__ word, const bf16* __restrict__ pos,
                 const bf16* __restrict__ type0) {
  __shared__ float red[kRowThreads / 32];
  const int r = blockIdx.x;
  const int nvec = H / 8;
  float vals[kMaxVec][8];
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int i = threadIdx.x + j * kRowThreads;
    if (i < nvec) {
      if (kEmbed) {
        float a[8], b[8], c[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(word + (size_t)tokens[r] * H) + i), a);
        unpack8(__ldg(reinterpret_cast<const uint4*>(pos + (size_t)positions[r] * H) + i), b);
        unpack8(__ldg(reinterpret_cast<const uint4*>(type0) + i), c);
#pragma unroll
        for (int k = 0; k < 8; ++k) vals[j][k] = a[k] + b[k] + c[k];
      } else {
        unpack8(reinterpret_cast<const uint4*>(x + (size_t)r * H)[i], vals[j]);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) sum += vals[j][k];
    }
  }
  const float mean = block_sum<kRowThreads>(sum, red) / (float)H;
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int i = threadIdx.x + j * kRowThreads;
    if (i < nvec) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float d = vals[j][k] - mean;
        sq += d * d;
      }
    }
  }
  const float inv = rsqrtf(block_sum<kRowThreads>(sq, red) / (float)H + eps);
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int i = threadIdx.x + j * kRowThreads;
    if (i < nvec) {
      float g[8], b[8], o[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(gamma) + i), g);
      unpack8(__ldg(reinterpret_cast<const uint4*>(beta) + i), b);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = (vals[j][k] - mean) * inv * g[k] + b[k];
      reinterpret_cast<uint4*>(out + (size_t)r * H)[i] = pack8(o);
    }
  }
}

// Warp-per-row LayerNorm for narrow rows (H <= 1024, e.g. the 768-wide encoder): no shared memory, no block barrier,
// four rows per 128-thread block; the block-per-row kernel above leaves a quarter of its threads idle at H = 768.
template <bool kEmbed>
__global__ void __launch_bounds__(128)
layernorm_warp_kernel(const bf16* __restrict__ x, const bf16* __restrict__ gamma, const bf16* __restrict__ beta,
                      bf16* __restrict__ out, int rows, int H, float eps, const int32_t* __restrict__ tokens,
                      const int32_t* __restrict__ positions, const bf16* __restrict__ word, const bf16* __restrict__ pos,
                      const bf16* __restrict__ type0) {
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const int nvec = H / 8;
  float vals[4][8];
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = lane + j * 32;
    if (i < nvec) {
      if (kEmbed) {
        float a[8], b[8], c[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(word + (size_t)tokens[r] * H) + i), a);
        unpack8(__ldg(reinterpret_cast<const uint4*>(pos + (size_t)positions[r] * H) + i), b);
        unpack8(__ldg(reinterpret_cast<const uint4*>(type0) + i), c);
#pragma unroll
        for (int k = 0; k < 8; ++k) vals[j][k] = a[k] + b[k] + c[k];
      } else {
        unpack8(reinterpret_cast<const uint4*>(x + (size_t)r * H)[i], vals[j]);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) sum += vals[j][k];
    }
  }
  const float mean = warp_sum(sum) / (float)H;
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (lane + j * 32 < nvec) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float d = vals[j][k] - mean;
        sq += d * d;
      }
    }
  }
  const float inv = rsqrtf(warp_sum(sq) / (float)H + eps);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = lane + j * 32;
    if (i < nvec) {
      float g[8], b[8], o[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(gamma) + i), g);
      unpack8(__ldg(reinterpret_cast<const uint4*>(beta) + i), b);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = (vals[j][k] - mean) * inv * g[k] + b[k];
      reinterpret_cast<uint4*>(out + (size_t)r * H)[i] = pack8(o);
    }
  }
}

__global__ void rope_table_kernel(const int32_t* __restrict__ positions, const float* __restrict__ inv_freq,
                                  float* __restrict__ cs, int T, int D) {
  const int half = D / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * half) return;
  const int t = i / half, k = i % half;
  float s, c;
  sincosf((float)positions[t] * inv_freq[k], &s, &c);
  cs[(size_t)t * D + k] = c;
  cs[(size_t)t * D + half + k] = s;
}

// One block per token. Pairs (i, i+D/2) of each q/k head are rotated; k,v rows go to the paged cache.
__global__ void __launch_bounds__(256)
rope_kv_write_kernel(bf16* __restrict__ qkv, const int32_t* __restrict__ positions,
                     const int32_t* __restrict__ slot_mapping, const float* __restrict__ inv_freq,
                     bf16* __restrict__ k_cache, bf16* __restrict__ v_cache, int Hq, int Hkv, int D, int page_size) {
  extern __shared__ float cs[];  // cos[D/2], sin[D/2]
  const int t = blockIdx.x;
  const int half = D / 2;
  const float p = (float)positions[t];
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    float s, c;
    sincosf(p * inv_freq[i], &s, &c);
    cs[i] = c;
    cs[half + i] = s;
  }
  __syncthreads();
  const int row_elems = (Hq + 2 * Hkv) * D;
  bf16* row = qkv + (size_t)t * row_elems;
  // rotate q and k heads: work items = (Hq+Hkv) * half/2 bf16x2 pairs
  const int pairs_per_head = half / 2;
  const int items = (Hq + Hkv) * pairs_per_head;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int h = it / pairs_per_head, j = (it % pairs_per_head) * 2;
    __nv_bfloat162* lo = reinterpret_cast<__nv_bfloat162*>(row + h * D + j);
    __nv_bfloat162* hi = reinterpret_cast<__nv_bfloat162*>(row + h * D + half + j);
    const float2 a = __bfloat1622float2(*lo), b = __bfloat1622float2(*hi);
    const float c0 = cs[j], c1 = cs[j + 1], s0 = cs[half + j], s1 = cs[half + j + 1];
    *lo = __floats2bfloat162_rn(rope_lo(a.x, b.x, c0, s0), rope_lo(a.y, b.y, c1, s1));
    *hi = __floats2bfloat162_rn(rope_hi(a.x, b.x, c0, s0), rope_hi(a.y, b.y, c1, s1));
  }
  const int slot = slot_mapping ? slot_mapping[t] : -1;
  if (slot < 0) return;
  __syncthreads();  // rotated k visible to the copy below
  const int page = slot / page_size, off = slot % page_size;
  const int vec_per_head = D / 8;
  const uint4* ksrc = reinterpret_cast<const uint4*>(row + Hq * D);
  const uint4* vsrc = reinterpret_cast<const uint4*>(row + (Hq + Hkv) * D);
  for (int it = threadIdx.x; it < Hkv * vec_per_head; it += blockDim.x) {
    const int h = it / vec_per_head, j = it % vec_per_head;
    const size_t dst = (((size_t)page * Hkv + h) * page_size + off) * vec_per_head + j;
    reinterpret_cast<uint4*>(k_cache)[dst] = ksrc[it];
    reinterpret_cast<uint4*>(v_cache)[dst] = vsrc[it];
  }
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

constexpr int kSampleChunk = 4096;  // vocabulary columns per stage-1 block

__device__ __forceinline__ void argmax_merge(float& best, int& bi, float ov, int oi) {
  if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
}
template <int NT>
__device__ __forceinline__ void block_argmax(float& best, int& bi, float* sval, int* sidx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    argmax_merge(best, bi, ov, oi);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sval[w] = best; sidx[w] = bi; }
  __syncthreads();
  if (w == 0) {
    best = (l < NT / 32) ? sval[l] : -INFINITY;
    bi = (l < NT / 32) ? sidx[l] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      argmax_merge(best, bi, ov, oi);
    }
  }
}

// ---- top-k / top-p (nucleus) filtering: stage 0 finds, per row, the logit threshold below which tokens are dropped ----
// float -> unsigned key that orders like the float (and back)
__device__ __forceinline__ uint32_t fkey(float x) {
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
constexpr int kThrThreads = 1024;
// block-wide sum of a (count, mass) pair; every thread gets the result; fixed summation tree -> deterministic
__device__ __forceinline__ void block_sum2(unsigned long long& cnt, unsigned long long& mass, unsigned long long* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    mass += __shfl_xor_sync(0xffffffffu, mass, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();  // sh reuse across calls
  if (l == 0) { sh[2 * w] = cnt; sh[2 * w + 1] = mass; }
  __syncthreads();
  cnt = sh[2 * l];
  mass = sh[2 * l + 1];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    mass += __shfl_xor_sync(0xffffffffu, mass, o);
  }
}
// One block per row. Result thr[b]: tokens with logit < thr[b] are excluded from sampling (-inf: keep all).
//   top-k : the largest threshold that keeps >= k tokens (ties at the k-th value are all kept)
//   top-p : applied to the top-k survivors: the smallest set of most-probable tokens whose softmax(logits/T) mass
//           reaches p of the survivors' mass
// Both are bisections over the ordered 32-bit key space (at most 32 passes over an L2-resident row each, usually ~18:
// the search stops as soon as one more step cannot change the kept set).  Probability mass is accumulated in 2^-32 fixed
// point so the result does not depend on summation order.
__global__ void __launch_bounds__(kThrThreads)
sample_threshold_kernel(const float* __restrict__ logits, int ldl, const float* __restrict__ temperature,
                        const int32_t* __restrict__ top_k, const float* __restrict__ top_p, float* __restrict__ thr, int V) {
  __shared__ unsigned long long sh[64];
  __shared__ float smax[32];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  const float temp = temperature ? temperature[b] : 0.f;
  const int k = top_k ? top_k[b] : 0;
  const float p = top_p ? top_p[b] : 1.f;
  const bool use_k = k > 0 && k < V, use_p = p > 0.f && p < 1.f;
  if (!(temp > 0.f) || (!use_k && !use_p)) {  // greedy rows and unfiltered rows
    if (threadIdx.x == 0) thr[b] = -INFINITY;
    return;
  }
  const float* row = logits + (size_t)b * ldl;
  // row max (for the softmax weights)
  float mx = -INFINITY;
  for (int v = threadIdx.x; v < V; v += kThrThreads) mx = fmaxf(mx, row[v]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = smax[threadIdx.x & 31];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const float inv_t = 1.f / temp;
  // (count, mass) of the tokens whose key is >= `key`
  auto tally = [&](uint32_t key, unsigned long long& cnt, unsigned long long& mass) {
    cnt = 0;
    mass = 0;
    for (int v = threadIdx.x; v < V; v += kThrThreads) {
      const float x = row[v];
      if (fkey(x) >= key) {
        cnt += 1;
        mass += (unsigned long long)(__expf((x - mx) * inv_t) * 4294967296.0f);
      }
    }
    block_sum2(cnt, mass, sh);
  };
  unsigned long long lo = 0, hi = 1ull << 32;  // keys; invariant: keeping [lo, ..) is enough, keeping [hi, ..) is not
  unsigned long long c_lo = V, c_hi = 0, m_lo = 0, c, m;
  if (use_k) {
    while (hi - lo > 1 && c_lo != (unsigned long long)k) {
      const unsigned long long mid = (lo + hi) >> 1;
      tally((uint32_t)mid, c, m);
      if (c >= (unsigned long long)k) { lo = mid; c_lo = c; } else { hi = mid; c_hi = c; }
    }
  }
  if (use_p) {
    tally((uint32_t)lo, c_lo, m_lo);  // mass of the top-k survivors (of the whole row without top-k)
    unsigned long long target = (unsigned long long)((double)p * (double)m_lo);
    if (target < 1) target = 1;
    hi = 1ull << 32;
    c_hi = 0;
    while (hi - lo > 1 && c_lo - c_hi > 1) {
      const unsigned long long mid = (lo + hi) >> 1;
      tally((uint32_t)mid, c, m);
      if (m >= target) { lo = mid; c_lo = c; } else { hi = mid; c_hi = c; }
    }
  }
  if (threadIdx.x == 0) thr[b] = lo == 0 ? -INFINITY : fkey_inv((uint32_t)lo);
}

// stage 1: grid (chunks, B): per-chunk argmax of logits/T + Gumbel noise -> part[b][chunk]
__global__ void __launch_bounds__(256)
sample_stage1_kernel(const float* __restrict__ logits, int ldl, const float* __restrict__ temperature,
                     const uint64_t* __restrict__ seed, const float* __restrict__ thr, float* __restrict__ part_val,
                     int* __restrict__ part_idx, int V) {
  __shared__ float sval[8];
  __shared__ int sidx[8];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y, chunk = blockIdx.x;
  const float* row = logits + (size_t)b * ldl;
  const float temp = temperature ? temperature[b] : 0.f;
  const bool greedy = !(temp > 0.f);
  const float inv_t = greedy ? 1.f : 1.f / temp;
  const uint64_t sd = seed ? seed[b] : 0;
  const float floor_logit = (thr && !greedy) ? thr[b] : -INFINITY;  // top-k / top-p survivors only
  float best = -INFINITY;
  int bi = 0x7fffffff;
  const int v_end = min(V, (chunk + 1) * kSampleChunk);
  for (int v = chunk * kSampleChunk + threadIdx.x; v < v_end; v += 256) {
    const float raw = row[v];
    if (raw < floor_logit) continue;
    float x = raw * inv_t;
    if (!greedy) {
      const uint64_t h = mix64(sd ^ (0xD1B54A32D192ED03ull * (uint64_t)(v + 1)));
      // 23 random bits + 0.5: every value is exact in fp32 and u stays strictly inside (0,1) — with 24 bits the
      // "+0.5f" rounds the top half of the range up and u == 1.0 made the noise +inf (a uniformly random token won
      // once per 2^24 draws, ~0.8 % of sampled tokens at a 128k vocabulary).  The inner log is the accurate one: the
      // fast-math log has an absolute error comparable to log(1 - 2^-24).
      const float u = ((float)(h >> 41) + 0.5f) * (1.0f / 8388608.0f);
      x += -__logf(-logf(u));
    }
    if (x > best || (x == best && v < bi)) { best = x; bi = v; }
  }
  block_argmax<256>(best, bi, sval, sidx);
  if (threadIdx.x == 0) {
    part_val[b * gridDim.x + chunk] = best;
    part_idx[b * gridDim.x + chunk] = bi;
  }
}
// stage 2: one warp per row merges the chunk winners (lowest index wins ties, as in stage 1)
__global__ void __launch_bounds__(32)
sample_stage2_kernel(const float* __restrict__ part_val, const int* __restrict__ part_idx, int32_t* __restrict__ out,
                     int chunks) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int c = threadIdx.x; c < chunks; c += 32) argmax_merge(best, bi, part_val[b * chunks + c], part_idx[b * chunks + c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    argmax_merge(best, bi, ov, oi);
  }
  if (threadIdx.x == 0) out[b] = bi < 0x7fffffff ? bi : 0;  // all-NaN row: emit a valid id, never an out-of-range one
}

// ---- OpenAI presence / frequency penalties: logits[b, tok] -= val for the (tok, val) entries of row b ----
// entries [pen_off[b], pen_off[b+1]) hold the DISTINCT tokens the sequence has generated so far with
// val = presence + frequency * count (computed on the host from the request's exact integer counts)
__global__ void __launch_bounds__(256)
apply_penalties_kernel(float* __restrict__ logits, int ldl, const int32_t* __restrict__ pen_off,
                       const int2* __restrict__ pen, int V) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  float* row = logits + (size_t)b * ldl;
  for (int i = pen_off[b] + threadIdx.x; i < pen_off[b + 1]; i += 256) {
    const int2 e = pen[i];
    if (e.x >= 0 && e.x < V) row[e.x] -= __int_as_float(e.y);  // distinct tokens: no two threads touch the same logit
  }
}

// ---- log-probabilities of the sampled token and of the width-1 most likely tokens (descending, lowest id on ties) ----
constexpr int kLpThreads = 1024;
__global__ void __launch_bounds__(kLpThreads)
logprob_topk_kernel(const float* __restrict__ logits, int ldl, int V, const int32_t* __restrict__ sampled,
                    const int32_t* __restrict__ width, int32_t* __restrict__ out_ids, float* __restrict__ out_lp,
                    int max_width) {
  __shared__ float sval[kLpThreads / 32];
  __shared__ int sidx[kLpThreads / 32];
  __shared__ float red[kLpThreads / 32];
  __shared__ float s_bv;
  __shared__ int s_bi;
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  const int w = min(width[b], max_width);
  if (w <= 0) return;
  const float* row = logits + (size_t)b * ldl;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int v = threadIdx.x; v < V; v += kLpThreads) argmax_merge(best, bi, row[v], v);
  block_argmax<kLpThreads>(best, bi, sval, sidx);
  if (threadIdx.x == 0) { s_bv = best; s_bi = bi; }
  __syncthreads();
  const float mx = s_bv;
  float se = 0.f;
  for (int v = threadIdx.x; v < V; v += kLpThreads) se += expf(row[v] - mx);
  se = block_sum<kLpThreads>(se, red);
  const float lse = mx + logf(se);
  int32_t* ids = out_ids + (size_t)b * max_width;
  float* lps = out_lp + (size_t)b * max_width;
  if (threadIdx.x == 0) {
    const int t = sampled[b];
    ids[0] = t;
    lps[0] = (t >= 0 && t < V) ? row[t] - lse : -INFINITY;
  }
  float pv = s_bv;   // the j-th most likely token in the order (value desc, id asc); j = 1 is the argmax found above
  int pi = s_bi;
  for (int j = 1; j < w; ++j) {
    if (threadIdx.x == 0) { ids[j] = pi; lps[j] = pv - lse; }
    if (j + 1 == w) break;
    best = -INFINITY;
    bi = 0x7fffffff;
    for (int v = threadIdx.x; v < V; v += kLpThreads) {
      const float x = row[v];
      if (x < pv || (x == pv && v > pi)) argmax_merge(best, bi, x, v);  // strictly after the previous pick
    }
    __syncthreads();  // sval/sidx reuse
    block_argmax<kLpThreads>(best, bi, sval, sidx);
    if (threadIdx.x == 0) { s_bv = best; s_bi = bi; }
    __syncthreads();
    pv = s_bv;
    pi = s_bi;
    if (pi == 0x7fffffff) {  // fewer than `w` finite logits in the row
      if (threadIdx.x == 0)
        for (int k = j + 1; k < w; ++k) { ids[k] = -1; lps[k] = -INFINITY; }
      break;
    }
  }
}

__global__ void __launch_bounds__(kRowThreads)
cls_pool_l2_kernel(const bf16* __restrict__ x, const int32_t* __restrict__ first_row, float* __restrict__ out, int H) {
  __shared__ float red[kRowThreads / 32];
  const int b = blockIdx.x;
  const bf16* row = x + (size_t)first_row[b] * H;
  float ss = 0.f;
  for (int i = threadIdx.x; i < H; i += kRowThreads) {
    const float v = __bfloat162float(row[i]);
    ss += v * v;
  }
  ss = block_sum<kRowThreads>(ss, red);
  const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  for (int i = threadIdx.x; i < H; i += kRowThreads) out[(size_t)b * H + i] = __bfloat162float(row[i]) * inv;
}

}  // namespace

cudaError_t embed_gather(cudaStream_t s, const int32_t* tokens, const bf16* table, bf16* x, int T, int H) {
  if (T <= 0) return cudaSuccess;
  if (H % 8) return cudaErrorInvalidValue;
  return launch_k(embed_gather_kernel, dim3(T), dim3(128), 0, s, true, tokens, table, x, H);
}
cudaError_t rmsnorm(cudaStream_t s, const bf16* x, const bf16* w, bf16* out, const int32_t* row_index, int rows, int H,
                    float eps) {
  if (rows <= 0) return cudaSuccess;
  if (H % 8) return cudaErrorInvalidValue;
  return launch_k(rmsnorm_kernel, dim3(rows), dim3(kRowThreads), 0, s, true, x, w, out, row_index, H, eps);
}
cudaError_t layernorm(cudaStream_t s, const bf16* x, const bf16* gamma, const bf16* beta, bf16* out, int rows, int H,
                      float eps) {
  if (rows <= 0) return cudaSuccess;
  if (H % 8 || H > kMaxVec * kRowThreads * 8) return cudaErrorInvalidValue;
  if (H <= 1024)
    layernorm_warp_kernel<false><<<(rows + 3) / 4, 128, 0, s>>>(x, gamma, beta, out, rows, H, eps, nullptr, nullptr, nullptr,
                                                                nullptr, nullptr);
  else
    layernorm_kernel<false><<<rows, kRowThreads, 0, s>>>(x, gamma, beta, out, H, eps, nullptr, nullptr, nullptr, nullptr, nullptr);
  return cudaGetLastError();
}
cudaError_t bert_embed_ln(cudaStream_t s, const int32_t* tokens, const int32_t* positions, const bf16* word,
                          const bf16* pos, const bf16* type0, const bf16* gamma, const bf16* beta, bf16* x, int T, int H,
                          float eps) {
  if (T <= 0) return cudaSuccess;
  if (H % 8 || H > kMaxVec * kRowThreads * 8) return cudaErrorInvalidValue;
  if (H <= 1024)
    layernorm_warp_kernel<true><<<(T + 3) / 4, 128, 0, s>>>(nullptr, gamma, beta, x, T, H, eps, tokens, positions, word, pos, type0);
  else
    layernorm_kernel<true><<<T, kRowThreads, 0, s>>>(nullptr, gamma, beta, x, H, eps, tokens, positions, word, pos, type0);
  return cudaGetLastError();
}
cudaError_t rope_table(cudaStream_t s, const int32_t* positions, const float* inv_freq, float* cs, int T, int D) {
  if (T <= 0) return cudaSuccess;
  const int n = T * (D / 2);
  rope_table_kernel<<<(n + 255) / 256, 256, 0, s>>>(positions, inv_freq, cs, T, D);
  return cudaGetLastError();
}

cudaError_t rope_kv_write(cudaStream_t s, bf16* qkv, const int32_t* positions, const int32_t* slot_mapping,
                          const float* inv_freq, bf16* k_cache, bf16* v_cache, int T, int Hq, int Hkv, int D,
                          int page_size) {
  if (T <= 0) return cudaSuccess;
  if (D % 8) return cudaErrorInvalidValue;
  rope_kv_write_kernel<<<T, 256, D * sizeof(float), s>>>(qkv, positions, slot_mapping, inv_freq, k_cache, v_cache, Hq,
                                                         Hkv, D, page_size);
  return cudaGetLastError();
}
size_t sample_scratch_bytes(int B, int V) { return (size_t)B * ((V + kSampleChunk - 1) / kSampleChunk) * 8 + (size_t)B * 4; }

cudaError_t sample_tokens(cudaStream_t s, const float* logits, int ldl, const float* temperature, const uint64_t* seed,
                          int32_t* out, int B, int V, void* scratch, const int32_t* top_k, const float* top_p) {
  if (B <= 0) return cudaSuccess;
  const int chunks = (V + kSampleChunk - 1) / kSampleChunk;
  float* pv = static_cast<float*>(scratch);
  int* pi = reinterpret_cast<int*>(pv + (size_t)B * chunks);
  float* thr = nullptr;
  cudaError_t e;
  if (temperature && (top_k || top_p)) {  // rows that ask for neither leave the kernel at once
    thr = reinterpret_cast<float*>(pi + (size_t)B * chunks);
    e = launch_k(sample_threshold_kernel, dim3(B), dim3(kThrThreads), 0, s, true, logits, ldl, temperature, top_k, top_p,
                 thr, V);
    if (e != cudaSuccess) return e;
  }
  e = launch_k(sample_stage1_kernel, dim3(chunks, B), dim3(256), 0, s, true, logits, ldl, temperature, seed,
               (const float*)thr, pv, pi, V);
  if (e != cudaSuccess) return e;
  return launch_k(sample_stage2_kernel, dim3(B), dim3(32), 0, s, true, (const float*)pv, (const int*)pi, out, chunks);
}
cudaError_t apply_penalties(cudaStream_t s, float* logits, int ldl, const int32_t* pen_off, const void* pen, int B, int V) {
  if (B <= 0) return cudaSuccess;
  return launch_k(apply_penalties_kernel, dim3(B), dim3(256), 0, s, true, logits, ldl, pen_off, (const int2*)pen, V);
}
cudaError_t logprob_topk(cudaStream_t s, const float* logits, int ldl, int V, const int32_t* sampled, const int32_t* width,
                         int32_t* out_ids, float* out_lp, int B, int max_width) {
  if (B <= 0) return cudaSuccess;
  return launch_k(logprob_topk_kernel, dim3(B), dim3(kLpThreads), 0, s, true, logits, ldl, V, sampled, width, out_ids,
                  out_lp, max_width);
}
cudaError_t cls_pool_l2(cudaStream_t s, const bf16* x, const int32_t* first_row, float* out, int B, int H) {
  if (B <= 0) return cudaSuccess;
  cls_pool_l2_kernel<<<B, kRowThreads, 0, s>>>(x, first_row, out, H);
  return cudaGetLastError();
}

}  // namespace hb
