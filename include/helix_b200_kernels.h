/* helix_b200_kernels.h — kernel-level C ABI (plain device pointers + sizes, stream 0 unless given).
 * These are the individual sm_100a ops K1–K11 of SURVEY.md §2.3 that the engine composes; they are
 * exported so parity tests (and a host shim that wants a single op) can drive each kernel in
 * isolation.  All return a cudaError_t value (0 = ok) or -1 for an invalid argument; launches are
 * asynchronous on the legacy default stream and the caller synchronises.
 */
#ifndef HELIX_B200_KERNELS_H_
#define HELIX_B200_KERNELS_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

int hbk_init(void);
const char* hbk_last_error(void);

/* C[M,N] = A[M,K] . W[N,K]^T, bf16 in; epi = enum hb::Epi (0 none,1 bias,2 bias+gelu,3 resid,4 bias+resid,
 * 5 swiglu (W packed [128 gate|128 up] per 256 rows, C has N/2 columns), 6 fp32 out). block_n 0 = auto. */
int hbk_gemm(const void* A, int lda, const void* W, int ldw, void* C, int ldc, const void* R, int ldr, const void* bias,
             int M, int N, int K, int epi, int block_n);
/* decode-step GEMM (M<=256): swap-AB + stream-K partial slabs summed in fixed order -> out fp32 [M,N] */
int hbk_gemm_skinny(const void* X, int ldx, const void* W, int ldw, float* out_f32, int ldo, int M, int N, int K);
/* the same GEMM with the fused tile finisher (last-arriving CTA of each output tile sums the slabs in slab order and
 * writes out[m, n]); launched `repeats` times back to back: the arrival counters must return to zero by themselves */
int hbk_gemm_skinny_finish(const void* X, int ldx, const void* W, int ldw, float* out_f32, int ldo, int M, int N, int K,
                           int repeats);
int hbk_gemm_naive(const void* A, int lda, const void* W, int ldw, void* C_f32, int ldc, int M, int N, int K);

int hbk_embed_gather(const int32_t* tokens, const void* table, void* x, int T, int H);
int hbk_bert_embed_ln(const int32_t* tokens, const int32_t* positions, const void* word, const void* pos,
                      const void* type0, const void* gamma, const void* beta, void* x, int T, int H, float eps);
int hbk_rmsnorm(const void* x, const void* w, void* out, const int32_t* row_index, int rows, int H, float eps);
int hbk_layernorm(const void* x, const void* gamma, const void* beta, void* out, int rows, int H, float eps);
int hbk_rope_kv_write(void* qkv, const int32_t* positions, const int32_t* slot_mapping, const float* inv_freq,
                      void* k_cache, void* v_cache, int T, int Hq, int Hkv, int D, int page_size);
/* QKV projection with RoPE + KV scatter in the GEMM epilogue (EPI_ROPE): qkv[T, (Hq+2Hkv)*D] = A[T,K] * W^T (+ bias), q / k heads
   rotated, k / v rows also written to the paged cache — must equal hbk_gemm followed by hbk_rope_kv_write bit for bit */
int hbk_gemm_qkv_rope(const void* A, int lda, const void* W, int ldw, void* qkv, const void* bias, const int32_t* positions,
                      const int32_t* slot_mapping, const float* inv_freq, void* k_cache, void* v_cache, int T, int K,
                      int Hq, int Hkv, int D, int page_size);
int hbk_sample(const float* logits, int ldl, const float* temperature, const uint64_t* seed, int32_t* out, int B, int V);
/* as hbk_sample, restricted per row to the top_k / top_p survivors (either pointer may be NULL) */
int hbk_sample_filtered(const float* logits, int ldl, const float* temperature, const uint64_t* seed, const int32_t* top_k,
                        const float* top_p, int32_t* out, int B, int V);
/* OpenAI presence / frequency penalties: logits[b, tok] -= val over row b's entries {int32 tok; float val}
 * [pen_off[b], pen_off[b+1]) (distinct tokens per row; val = presence + frequency * count, computed by the caller) */
int hbk_apply_penalties(float* logits, int ldl, const int32_t* pen_off, const void* pen_entries, int B, int V);
/* rows with width[b] > 0: out_ids/out_lp[b][0] = sampled[b] and its log-softmax; [1..width[b]) = the most likely tokens in
 * descending order (lowest id first on ties).  out_* are [B][max_width]. */
int hbk_logprob_topk(const float* logits, int ldl, int V, const int32_t* sampled, const int32_t* width, int32_t* out_ids,
                     float* out_lp, int B, int max_width);
int hbk_cls_pool_l2(const void* x, const int32_t* first_row, float* out, int B, int H);

int hbk_attn_prefill(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo,
                     const int32_t* cu_seqlens, int B, int T, int max_seqlen, int Hq, int Hkv, int D, int causal,
                     float scale);
/* chunked prefill: the cu_seqlens[b+1]-cu_seqlens[b] q rows of sequence b are the LAST positions of a kv_lens[b]-long
 * sequence whose K/V (this chunk's included) are in the paged pool plane [num_pages][Hkv][64][D] (page size 64). */
int hbk_attn_prefill_paged(const void* q, int ldq, const void* k_cache, const void* v_cache, const int32_t* page_table,
                           int max_pages, const int32_t* kv_lens, void* out, int ldo, const int32_t* cu_seqlens, int B,
                           int T, int max_q_len, int Hq, int Hkv, int D, int causal, float scale, int num_pages);
int hbk_attn_naive(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, float* out_f32, int ldo,
                   const int32_t* cu_seqlens, int B, int T, int max_seqlen, int Hq, int Hkv, int D, int causal,
                   float scale);
int hbk_attn_decode(const void* q, int ldq, const void* k_cache, const void* v_cache, const int32_t* page_table,
                    int max_pages, const int32_t* ctx_lens, void* out, int ldo, float* workspace, int B, int Hq, int Hkv,
                    int D, int page_size, int num_splits, float scale, int num_pages);
size_t hbk_attn_decode_workspace_floats(int B, int Hq, int D, int num_splits);

#ifdef __cplusplus
}
#endif
#endif
