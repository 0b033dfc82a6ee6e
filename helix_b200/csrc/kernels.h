// Internal C++ interface to the sm_100a kernels (the C ABI in include/helix_b200.h sits on top).
// All pointers are device pointers; every launcher is asynchronous on `stream` and returns the
// launch status.  Shapes follow the reference backends' conventions (SURVEY.md §2.3, K1–K11).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace hb {

typedef __nv_bfloat16 bf16;

// Sets the opt-in shared-memory attribute of every kernel instantiation; call once per process/device.
cudaError_t kernels_init();
cudaError_t gemm_init();
cudaError_t attn_prefill_init();
cudaError_t attn_decode_init();

// ---- GEMM: C[M,N] = A[M,K] · W[N,K]^T (both K-major, bf16 in, fp32 accumulate in TMEM) ----
enum Epi : int {
  EPI_NONE = 0,        // C = acc                              (bf16)
  EPI_BIAS = 1,        // C = acc + bias[n]                    (bf16)
  EPI_BIAS_GELU = 2,   // C = gelu_erf(acc + bias[n])          (bf16)
  EPI_RESID = 3,       // C = acc + R[m,n]                     (bf16, R may alias C)
  EPI_BIAS_RESID = 4,  // C = acc + bias[n] + R[m,n]           (bf16)
  EPI_SWIGLU = 5,      // W rows packed per 256-row tile as [128 gate | 128 up];
                       // C[m, tile*128+j] = silu(gate_j) * up_j   (bf16, ldc counts N/2 columns)
  EPI_F32 = 6,         // C = acc                              (fp32)
  EPI_ROPE = 7,        // fused QKV projection of a decoder: C = bf16(acc [+ bias]) with the q and k heads rotated
                       // (RopeEpi) and the k / v rows also scattered into the paged KV cache
};

// What EPI_ROPE needs beyond the GEMM itself (the job of rope_kv_write, done on the accumulators' way out).
struct RopeEpi {
  bf16* out = nullptr;             // == GemmArgs::C, written with plain stores (row stride ldc elements)
  int ldc = 0;
  const float* cs = nullptr;       // [M][D]: cos(pos*inv_freq[i]) for i < D/2, then sin (rope_table)
  const int32_t* slots = nullptr;  // slot_mapping as in rope_kv_write; null or < 0: no cache write for the row
  bf16* k_cache = nullptr;
  bf16* v_cache = nullptr;
  int Hq = 0, Hkv = 0, D = 0, page_size = 0;
};

struct GemmArgs {
  const bf16* A;
  int lda;
  const bf16* W;
  int ldw;
  void* C;
  int ldc;
  const bf16* R;
  int ldr;
  const bf16* bias;
  int M, N, K;
  Epi epi;
  int block_n;  // 0 = choose
  RopeEpi rope{};  // EPI_ROPE only
};
cudaError_t gemm_bf16_tn(cudaStream_t stream, const GemmArgs& a);
// Debug/test-only CUDA-core GEMM (same contract, EPI_NONE/EPI_F32 only); used by tests as an on-device checker.
cudaError_t gemm_naive_check(cudaStream_t stream, const GemmArgs& a);


// ---- HBM-bound row kernels (K1, K2, K4, K10 sampling, K11) ----
// x[t,:] = table[tokens[t],:]
cudaError_t embed_gather(cudaStream_t s, const int32_t* tokens, const bf16* table, bf16* x, int T, int H);
// BERT: x[t,:] = LayerNorm(word[tok[t]] + pos[positions[t]] + type[0]) * g + b
cudaError_t bert_embed_ln(cudaStream_t s, const int32_t* tokens, const int32_t* positions, const bf16* word,
                          const bf16* pos, const bf16* type0, const bf16* gamma, const bf16* beta, bf16* x, int T, int H,
                          float eps);
// out[r,:] = x[row_index ? row_index[r] : r, :] * rsqrt(mean(x^2)+eps) * w    (fp32 math, one bf16 rounding)
cudaError_t rmsnorm(cudaStream_t s, const bf16* x, const bf16* w, bf16* out, const int32_t* row_index, int rows, int H,
                    float eps);
// out[r,:] = (x[r,:]-mean)/sqrt(var+eps)*g + b   (in-place allowed)
cudaError_t layernorm(cudaStream_t s, const bf16* x, const bf16* gamma, const bf16* beta, bf16* out, int rows, int H,
                      float eps);
// In-place rotary embedding (HF rotate_half convention) on the q and k column blocks of the fused
// qkv[T, (Hq+2*Hkv)*D] buffer, then scatter of the rotated k and of v into the paged KV cache:
//   slot = slot_mapping[t] = page*page_size + offset;  cache layout [page][Hkv][page_size][D].
// slot < 0 skips the cache write (embedding-only models never call this).
cudaError_t rope_kv_write(cudaStream_t s, bf16* qkv, const int32_t* positions, const int32_t* slot_mapping,
                          const float* inv_freq, bf16* k_cache, bf16* v_cache, int T, int Hq, int Hkv, int D,
                          int page_size);
// cs[t][0..D/2) = cos(positions[t] * inv_freq[i]), cs[t][D/2..D) = sin(...): the table EPI_ROPE reads (once per step,
// shared by every layer).
cudaError_t rope_table(cudaStream_t s, const int32_t* positions, const float* inv_freq, float* cs, int T, int D);
// Greedy / Gumbel-max sampling over fp32 logits[B,V]: out[b] = argmax_v(logits[b,v]/temp[b] + g(seed[b],v)),
// temp[b] <= 0 -> pure argmax (lowest index wins ties).
cudaError_t sample_tokens(cudaStream_t s, const float* logits, int ldl, const float* temperature, const uint64_t* seed,
                          int32_t* out, int B, int V, void* scratch, const int32_t* top_k = nullptr,
                          const float* top_p = nullptr);  // per-row top-k (<=0 off) / top-p (outside (0,1) off), sampled rows only
size_t sample_scratch_bytes(int B, int V);  // two-stage argmax partials
// logits[b, tok] -= val for the entries {int32 tok; float val} [pen_off[b], pen_off[b+1]) of row b (distinct tokens per row)
cudaError_t apply_penalties(cudaStream_t s, float* logits, int ldl, const int32_t* pen_off, const void* pen, int B, int V);
// rows with width[b] > 0: out_ids/out_lp[b][0] = sampled token and its log-softmax, [1..width) = most likely tokens (desc)
cudaError_t logprob_topk(cudaStream_t s, const float* logits, int ldl, int V, const int32_t* sampled, const int32_t* width,
                         int32_t* out_ids, float* out_lp, int B, int max_width);
// out[b,:] = l2normalize(x[first_row[b], :]) as fp32 (CLS pooling for bge-style encoders)
cudaError_t cls_pool_l2(cudaStream_t s, const bf16* x, const int32_t* first_row, float* out, int B, int H);


// ---- K5: varlen flash attention over contiguous q/k/v rows (whole-prompt prefill, BERT encoder) ----
struct AttnPrefillArgs {
  const bf16* q; int ldq;   // [T, >=Hq*D], head h at column h*D
  const bf16* k; int ldk;   // [T, >=Hkv*D]
  const bf16* v; int ldv;
  bf16* out; int ldo;       // [T, Hq*D]
  const int32_t* cu_seqlens;  // [B+1] token offsets
  int B, T, max_seqlen;
  int Hq, Hkv, D;
  int causal;
  float scale;              // 1/sqrt(D)
  // chunked prefill (optional; k_cache != nullptr selects it): the q rows are the last rows of kv_lens[b]-long sequences
  // whose K/V are read from the paged pool instead of k/v
  const bf16* k_cache = nullptr;   // [num_pages][Hkv][64][D] plane of one layer
  const bf16* v_cache = nullptr;
  const int32_t* page_table = nullptr;  // [B, max_pages]
  const int32_t* kv_lens = nullptr;     // [B] kv length including this step's q rows
  int max_pages = 0, num_pages = 0, page_size = 64;
};
cudaError_t attn_prefill(cudaStream_t stream, const AttnPrefillArgs& a);
// test-only CUDA-core checker, fp32 output [T, ldo]
cudaError_t attn_naive_check(cudaStream_t stream, const AttnPrefillArgs& a, float* out_f32);

// Stream hand-over between consecutive HBM-streaming kernels of a decode step (see ptx.cuh sig_*): `wait` is the counter
// the PREVIOUS streaming kernel bumps once per CTA when its last load is issued (`wait_count` = its CTA count), `done`
// is this kernel's own counter, `bank_bytes` how much of its own stream the kernel may prefetch into L2 once the
// predecessor's loads have ended.  All-zero = feature off.
// Data dependency of a decode-step kernel on its predecessor as a release/acquire counter (ptx.cuh dep_*): `wait` is the
// predecessor's counter, `wait_count` how many of its CTAs report, `done` this kernel's own counter.  wait == nullptr:
// the kernel uses griddepcontrol.wait (every standalone / prefill use).
struct DepSig {
  const int* wait = nullptr;
  int wait_count = 0;
  int* done = nullptr;
};
struct StreamSig {
  const int* wait = nullptr;
  int wait_count = 0;
  int* done = nullptr;
  size_t bank_bytes = 0;
  unsigned long long* trace = nullptr;  // 16 event slots of this launch (HB_DEC_TRACE), written by CTA 0
  DepSig dep;                           // data dependency on the previous kernel of the step (optional)
};

// ---- K6: paged-KV decode attention (one query token per sequence), split-KV + combine ----
// Optional prologue of the decode-attention kernel (one split only): each CTA (sequence b, kv head) is the ONLY consumer of
// that sequence's q heads of the group and of its new K/V row, so it builds them itself from the QKV GEMM's slabs — slab sum
// (+ bias), bf16 rounding, RoPE, K/V written into the paged cache, q staged straight into shared memory — and the separate
// RoPE / KV-write kernel (and its kernel boundary) disappears from the layer.  Same arithmetic as dec_qkv_rope_kvwrite.
struct DecodeRope {
  const float* ws = nullptr;        // QKV GEMM slabs [seg][M][N]; nullptr = q is read from AttnDecodeArgs::q as usual
  const uint8_t* segs = nullptr;    // slabs per 128-column tile
  int M = 0, N = 0;                 // batch rows, (Hq + 2 Hkv) * D
  const bf16* bias = nullptr;       // [N] or null
  const int32_t* positions = nullptr;
  const int32_t* slots = nullptr;
  const float* inv_freq = nullptr;
  bf16* k_cache = nullptr;          // writable views of the layer's planes
  bf16* v_cache = nullptr;
};
struct AttnDecodeArgs {
  const bf16* q; int ldq;          // [B, >=Hq*D] (post-RoPE)
  const bf16* k_cache;             // [num_pages][Hkv][page_size][D]
  const bf16* v_cache;
  const int32_t* page_table;       // [B, max_pages]
  int max_pages;
  const int32_t* ctx_lens;         // [B] kv length including the current token
  bf16* out; int ldo;              // [B, Hq*D]
  float* workspace;                // >= B*Hq*num_splits*(D+2) floats
  int B, Hq, Hkv, D, page_size, num_splits;
  float scale;
  int num_pages;                   // pages in the pool (TMA tensor-map extent)
  StreamSig sig;                   // HBM hand-over with the neighbouring streaming kernels (optional)
  DecodeRope rope;                 // build q / K / V from the QKV GEMM's slabs inside the kernel (num_splits == 1 only)
};
cudaError_t attn_decode(cudaStream_t stream, const AttnDecodeArgs& a);
size_t attn_decode_workspace_floats(int B, int Hq, int D, int num_splits);


// ---- decode-step GEMM (M <= 256): swap-AB + deterministic stream-K, fp32 partial slabs ----
struct SkinnyPlan {
  int grid;                  // CTAs (== SMs unless the problem is tiny)
  int n_tiles;               // ceil(N/128)
  int max_segs;              // max #slabs any tile is split into
  const uint8_t* seg_count;  // device [n_tiles]: slabs of each 128-column tile
};
// Tile finisher of the decode GEMM: the LAST CTA that contributes a slab to an output tile (one counter per tile) sums
// the tile's slabs in slab order — bit-deterministic whoever finishes — and applies the op that follows the projection
// in the decoder layer, so no separate row kernel (and no kernel boundary) sits between two projections.
enum SkinnyEpiMode : int {
  SK_SLABS = 0,       // no finisher: slabs only (legacy path, consumed by the dec_* row kernels)
  SK_F32 = 1,         // out[m, n] = rstd[m] * sum                                   (LM head -> logits)
  SK_QKV_ROPE = 2,    // bf16(rstd[m] * sum) -> RoPE on q / k heads -> q to qkv_out, k / v into the paged cache
  SK_RESID_NORM = 3,  // x = bf16(x + sum); xg = bf16(x * gain[n]); ss_out[tile][m] = sum_n x^2 over the tile's columns
  SK_SWIGLU = 4,      // tiles (2t, 2t+1) = gate / up rows of block t:  h[m, t*128+j] = silu(rstd*gate) * (rstd*up)
};
constexpr int kSkinnySsStride = 256;  // row stride of the per-tile sum-of-squares partials ([tile][m], m < 256)
struct SkinnyEpi {
  int mode = SK_SLABS;
  int* tile_cnt = nullptr;          // per-engine arrival counters (zero between launches; the finisher re-zeroes its own)
  // rstd[m] = rsqrt(sum_t ss_in[t * 256 + m] / norm_h + eps): RMSNorm statistics of the INPUT rows, carried as per-tile
  // partials by the SK_RESID_NORM finisher of the previous projection (the GEMM is linear in the row factor); null: 1
  const float* ss_in = nullptr;
  int ss_tiles = 0, norm_h = 0;
  float eps = 0.f;
  float* out_f32 = nullptr; int ldo = 0;                                        // SK_F32
  bf16* qkv_out = nullptr; const int32_t* positions = nullptr; const int32_t* slots = nullptr;  // SK_QKV_ROPE
  const float* inv_freq = nullptr; bf16* k_cache = nullptr; bf16* v_cache = nullptr;
  int Hq = 0, Hkv = 0, D = 0, page_size = 64;
  bf16* x = nullptr; const bf16* gain = nullptr; bf16* xg = nullptr; float* ss_out = nullptr;   // SK_RESID_NORM
  bf16* h = nullptr; int F = 0;                                                  // SK_SWIGLU
};
// x = table[token]; xg = bf16(x * gain); ss_out[tile][m] = per-128-column sums of x^2 (what SK_RESID_NORM leaves behind)
cudaError_t dec_embed_prep(cudaStream_t s, const int32_t* tokens, const bf16* table, const bf16* gain, bf16* x, bf16* xg,
                           float* ss_out, int M, int H);
cudaError_t gemm_skinny_init();
cudaError_t gemm_skinny_plan(int N, int K, SkinnyPlan* out);  // cached per (device, N, K)
size_t gemm_skinny_ws_floats(const SkinnyPlan& p, int M, int N);
int gemm_skinny_max_segs(int N, int K, int sms);  // pure host arithmetic (memory estimates)
// ws[seg][m][n] (seg < seg_count[n/128]) = partial sums of X[M,K] . W[N,K]^T
cudaError_t gemm_skinny(cudaStream_t stream, const SkinnyPlan& plan, const bf16* X, int ldx, const bf16* W, int ldw,
                        float* ws, int M, int N, int K, const StreamSig* sig = nullptr, const SkinnyEpi* epi = nullptr);
// consumers of the slabs (each sums the slabs in fixed order, then applies its fused epilogue)
cudaError_t dec_sum_slabs(cudaStream_t s, const float* ws, const SkinnyPlan& p, float* out, int ldo, int M, int N);
// the dec_* kernels below take an optional DepSig; *ctas (when given) receives the number of CTAs that will report
cudaError_t dec_qkv_rope_kvwrite(cudaStream_t s, const float* ws, const SkinnyPlan& p, bf16* qkv_out,
                                 const int32_t* positions, const int32_t* slot_mapping, const float* inv_freq,
                                 bf16* k_cache, bf16* v_cache, int M, int Hq, int Hkv, int D, int page_size,
                                 const DepSig* dep = nullptr, const bf16* qkv_bias = nullptr);
int dec_qkv_rope_ctas(int M, int Hq, int Hkv);
int dec_swiglu_ctas(int M, int F);
// x = bf16(x + sum); xn = rmsnorm(x) * w (w may be null: residual update only)
cudaError_t dec_resid_rmsnorm(cudaStream_t s, const float* ws, const SkinnyPlan& p, bf16* x, const bf16* w, bf16* xn,
                              int M, int H, float eps, const DepSig* dep = nullptr);
// h[m, t*128+j] = silu(sum[m, t*256+j]) * sum[m, t*256+128+j]
cudaError_t dec_swiglu(cudaStream_t s, const float* ws, const SkinnyPlan& p, bf16* h, int M, int F, const DepSig* dep = nullptr);

}  // namespace hb
