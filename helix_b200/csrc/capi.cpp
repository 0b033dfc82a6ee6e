// C ABI (include/helix_b200.h, include/helix_b200_kernels.h) over hb::Engine and the kernel launchers.
#include <charconv>
#include <string.h>

#include <string>

#include "../../include/helix_b200.h"
#include "../../include/helix_b200_kernels.h"
#include "engine.h"
#include "gguf.h"
#include "tma_host.h"

using hb::Engine;

struct hb_engine {
  Engine impl;
  explicit hb_engine(const hb_engine_cfg& c) : impl(c) {}
};

namespace {
thread_local std::string g_last_create_error;
thread_local std::string g_kernel_error;
int kret(cudaError_t e) {
  if (e != cudaSuccess) g_kernel_error = std::string(cudaGetErrorString(e)) + " [" + hb::tmap_last_error() + "]";
  return (int)e;
}
}  // namespace

extern "C" {

int hb_abi_version(void) { return HB_ABI_VERSION; }

int hb_engine_create(const hb_engine_cfg* cfg, hb_engine** out) {
  if (!cfg || !out) {
    g_last_create_error = "null argument";
    return HB_ERR_INVALID;
  }
  hb_engine* e = new hb_engine(*cfg);
  int rc = e->impl.init();
  if (rc != HB_OK) {
    g_last_create_error = e->impl.last_error();
    delete e;
    *out = nullptr;
    return rc;
  }
  *out = e;
  return HB_OK;
}
void hb_engine_destroy(hb_engine* e) { delete e; }
const char* hb_last_error(hb_engine* e) { return e ? e->impl.last_error() : g_last_create_error.c_str(); }

int hb_model_load_begin(hb_engine* e, const hb_model_desc* d) { return (e && d) ? e->impl.load_begin(*d) : HB_ERR_INVALID; }
int hb_model_tensor_set(hb_engine* e, const char* name, const void* host, size_t n) {
  return (e && name && host) ? e->impl.tensor_set(name, host, n) : HB_ERR_INVALID;
}
int hb_model_load_finish(hb_engine* e) { return e ? e->impl.load_finish() : HB_ERR_INVALID; }
int hb_model_load_random(hb_engine* e, const hb_model_desc* d, uint64_t seed) {
  return (e && d) ? e->impl.load_random(*d, seed) : HB_ERR_INVALID;
}
int hb_model_load_gguf(hb_engine* e, const char* path) { return e ? e->impl.load_gguf(path) : HB_ERR_INVALID; }
int hb_gguf_describe(const char* path, hb_model_desc* desc) {
  if (!path || !desc) return HB_ERR_INVALID;
  hb::GgufFile g;
  std::string err;
  if (!g.open(path, &err) || !g.describe(desc, &err)) {
    g_last_create_error = "gguf: " + err;
    return HB_ERR_INVALID;
  }
  return HB_OK;
}
int hb_gguf_read_tensor(const char* path, const char* hf_name, float* out, size_t cap, size_t* rows, size_t* cols) {
  if (!path || !hf_name || !rows || !cols) return HB_ERR_INVALID;
  hb::GgufFile g;
  std::string err;
  hb_model_desc d;
  if (!g.open(path, &err) || !g.describe(&d, &err)) {
    g_last_create_error = "gguf: " + err;
    return HB_ERR_INVALID;
  }
  for (const auto& kv : g.tensors()) {
    if (hb::gguf_to_hf_name(kv.first) != hf_name) continue;
    std::vector<float> f;
    if (!g.read_f32(kv.first, &f, rows, cols)) return HB_ERR_INVALID;
    const bool permute = g.str("general.architecture") == "llama";
    if (permute && strstr(hf_name, "self_attn.q_proj.weight")) hb::gguf_unpermute_rows(f, *rows, *cols, d.heads);
    if (permute && strstr(hf_name, "self_attn.k_proj.weight")) hb::gguf_unpermute_rows(f, *rows, *cols, d.kv_heads);
    if (!out || cap < f.size()) return HB_ERR_BUSY;  // *rows x *cols floats needed
    memcpy(out, f.data(), f.size() * 4);
    return HB_OK;
  }
  return HB_ERR_NOT_FOUND;
}
int hb_model_weights_arena(hb_engine* e, void** p, size_t* b) { return e ? e->impl.weights_arena(p, b) : HB_ERR_INVALID; }
int hb_memory_estimate(const hb_model_desc* d, const hb_engine_cfg* c, uint64_t* w, uint64_t* kv, uint64_t* ws) {
  if (!d || !c) return HB_ERR_INVALID;
  if (!hb::validate_desc(*d).empty()) return HB_ERR_INVALID;
  Engine::estimate(*d, *c, w, kv, ws);
  return HB_OK;
}

int hb_engine_start(hb_engine* e) { return e ? e->impl.start() : HB_ERR_INVALID; }
int hb_engine_stop(hb_engine* e) { return e ? e->impl.stop() : HB_ERR_INVALID; }
int hb_engine_set_mixed(hb_engine* e, int32_t on, int32_t tokens) { return e ? e->impl.set_mixed(on, tokens) : HB_ERR_INVALID; }
int hb_step(hb_engine* e, int* did) { return e ? e->impl.step(did) : HB_ERR_INVALID; }
int hb_submit(hb_engine* e, const int32_t* t, int32_t n, const hb_sampling* sp, uint64_t* id) {
  return e ? e->impl.submit(t, n, sp, id) : HB_ERR_INVALID;
}
int hb_poll(hb_engine* e, uint64_t id, int32_t* out, int32_t cap, int32_t* n, int32_t* fin) {
  return e ? e->impl.poll(id, out, cap, n, fin) : HB_ERR_INVALID;
}
int hb_wait(hb_engine* e, uint64_t id, int32_t ms) { return e ? e->impl.wait(id, ms) : HB_ERR_INVALID; }
int hb_cancel(hb_engine* e, uint64_t id) { return e ? e->impl.cancel(id) : HB_ERR_INVALID; }
int hb_release(hb_engine* e, uint64_t id) { return e ? e->impl.release(id) : HB_ERR_INVALID; }
int hb_captured_logits(hb_engine* e, uint64_t id, int32_t which, float* out, size_t cap, int32_t* rows) {
  return e ? e->impl.captured(id, which, out, cap, rows) : HB_ERR_INVALID;
}
int hb_logprobs(hb_engine* e, uint64_t id, int32_t first_row, int32_t max_rows, int32_t* ids, float* lps, int32_t* rows,
                int32_t* width) {
  return e ? e->impl.logprobs(id, first_row, max_rows, ids, lps, rows, width) : HB_ERR_INVALID;
}
int hb_replica_unique_id(void* id) { return Engine::replica_unique_id(id); }
int hb_model_load_broadcast(hb_engine* e, const hb_model_desc* d, const void* id, int32_t rank, int32_t world, double* seconds) {
  return (e && d) ? e->impl.load_broadcast(*d, id, rank, world, seconds) : HB_ERR_INVALID;
}
int hb_embed(hb_engine* e, const int32_t* t, const int32_t* off, int32_t n, float* out) {
  return e ? e->impl.embed(t, off, n, out) : HB_ERR_INVALID;
}
int hb_get_stats(hb_engine* e, hb_stats* s) { return e ? e->impl.stats(s) : HB_ERR_INVALID; }
int hb_set_profile(hb_engine* e, int32_t on) { return e ? e->impl.set_profile(on != 0) : HB_ERR_INVALID; }

// ------------------------------------------------------------------ kernel-level ABI
int hbk_init(void) { return kret(hb::kernels_init()); }
const char* hbk_last_error(void) { return g_kernel_error.c_str(); }

int hbk_gemm(const void* A, int lda, const void* W, int ldw, void* C, int ldc, const void* R, int ldr, const void* bias,
             int M, int N, int K, int epi, int block_n) {
  hb::GemmArgs g{(const hb::bf16*)A, lda, (const hb::bf16*)W, ldw, C, ldc, (const hb::bf16*)R, ldr,
                 (const hb::bf16*)bias, M, N, K, (hb::Epi)epi, block_n};
  return kret(hb::gemm_bf16_tn(0, g));
}
int hbk_gemm_skinny(const void* X, int ldx, const void* W, int ldw, float* out, int ldo, int M, int N, int K) {
  hb::SkinnyPlan plan;
  cudaError_t e = hb::gemm_skinny_plan(N, K, &plan);
  if (e != cudaSuccess) return kret(e);
  float* ws = nullptr;
  if ((e = cudaMalloc(&ws, hb::gemm_skinny_ws_floats(plan, M, N) * 4)) != cudaSuccess) return kret(e);
  e = hb::gemm_skinny(0, plan, (const hb::bf16*)X, ldx, (const hb::bf16*)W, ldw, ws, M, N, K);
  if (e == cudaSuccess) e = hb::dec_sum_slabs(0, ws, plan, out, ldo, M, N);
  cudaError_t e2 = cudaDeviceSynchronize();
  cudaFree(ws);
  return kret(e != cudaSuccess ? e : e2);
}
int hbk_gemm_skinny_finish(const void* X, int ldx, const void* W, int ldw, float* out, int ldo, int M, int N, int K, int repeats) {
  hb::SkinnyPlan plan;
  cudaError_t e = hb::gemm_skinny_plan(N, K, &plan);
  if (e != cudaSuccess) return kret(e);
  float* ws = nullptr;
  int* cnt = nullptr;
  if ((e = cudaMalloc(&ws, hb::gemm_skinny_ws_floats(plan, M, N) * 4)) != cudaSuccess) return kret(e);
  if ((e = cudaMalloc(&cnt, (size_t)plan.n_tiles * 4)) != cudaSuccess) return kret(e);
  cudaMemset(cnt, 0, (size_t)plan.n_tiles * 4);
  hb::SkinnyEpi epi{};
  epi.mode = hb::SK_F32;
  epi.tile_cnt = cnt;
  epi.out_f32 = out;
  epi.ldo = ldo;
  for (int r = 0; r < repeats && e == cudaSuccess; ++r)  // the counters must come back to zero by themselves
    e = hb::gemm_skinny(0, plan, (const hb::bf16*)X, ldx, (const hb::bf16*)W, ldw, ws, M, N, K, nullptr, &epi);
  cudaError_t e2 = cudaDeviceSynchronize();
  cudaFree(ws);
  cudaFree(cnt);
  return kret(e != cudaSuccess ? e : e2);
}
int hbk_gemm_naive(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K) {
  hb::GemmArgs g{(const hb::bf16*)A, lda, (const hb::bf16*)W, ldw, C, ldc, nullptr, 0, nullptr, M, N, K, hb::EPI_F32, 0};
  return kret(hb::gemm_naive_check(0, g));
}
int hbk_embed_gather(const int32_t* tokens, const void* table, void* x, int T, int H) {
  return kret(hb::embed_gather(0, tokens, (const hb::bf16*)table, (hb::bf16*)x, T, H));
}
int hbk_bert_embed_ln(const int32_t* tokens, const int32_t* positions, const void* word, const void* pos,
                      const void* type0, const void* gamma, const void* beta, void* x, int T, int H, float eps) {
  return kret(hb::bert_embed_ln(0, tokens, positions, (const hb::bf16*)word, (const hb::bf16*)pos,
                                (const hb::bf16*)type0, (const hb::bf16*)gamma, (const hb::bf16*)beta, (hb::bf16*)x, T, H,
                                eps));
}
int hbk_rmsnorm(const void* x, const void* w, void* out, const int32_t* row_index, int rows, int H, float eps) {
  return kret(hb::rmsnorm(0, (const hb::bf16*)x, (const hb::bf16*)w, (hb::bf16*)out, row_index, rows, H, eps));
}
int hbk_layernorm(const void* x, const void* g, const void* b, void* out, int rows, int H, float eps) {
  return kret(hb::layernorm(0, (const hb::bf16*)x, (const hb::bf16*)g, (const hb::bf16*)b, (hb::bf16*)out, rows, H, eps));
}
int hbk_rope_kv_write(void* qkv, const int32_t* positions, const int32_t* slot_mapping, const float* inv_freq,
                      void* k_cache, void* v_cache, int T, int Hq, int Hkv, int D, int page_size) {
  return kret(hb::rope_kv_write(0, (hb::bf16*)qkv, positions, slot_mapping, inv_freq, (hb::bf16*)k_cache,
                                (hb::bf16*)v_cache, T, Hq, Hkv, D, page_size));
}
size_t hb_json_f32_array(const float* v, size_t n, char* out, size_t cap) {
  size_t len = 0;
  auto put = [&](const char* p, size_t k) {
    if (out && len + k <= cap) memcpy(out + len, p, k);
    len += k;
  };
  put("[", 1);
  char buf[32];
  for (size_t i = 0; i < n; ++i) {
    if (i) put(",", 1);
    const float x = v[i];
    if (!(x - x == 0.0f)) {  // NaN / inf have no JSON spelling
      put("null", 4);
      continue;
    }
    const auto r = std::to_chars(buf, buf + sizeof buf, x);  // shortest round-trip text
    put(buf, (size_t)(r.ptr - buf));
  }
  put("]", 1);
  return len;
}

int hbk_gemm_qkv_rope(const void* A, int lda, const void* W, int ldw, void* qkv, const void* bias, const int32_t* positions,
                      const int32_t* slot_mapping, const float* inv_freq, void* k_cache, void* v_cache, int T, int K,
                      int Hq, int Hkv, int D, int page_size) {
  const int N = (Hq + 2 * Hkv) * D;
  float* cs = nullptr;
  cudaError_t e = cudaMalloc(&cs, (size_t)T * D * 4);
  if (e != cudaSuccess) return kret(e);
  e = hb::rope_table(0, positions, inv_freq, cs, T, D);
  if (e == cudaSuccess) {
    hb::GemmArgs g{(const hb::bf16*)A, lda, (const hb::bf16*)W, ldw, qkv, N, nullptr, 0, (const hb::bf16*)bias, T, N, K, hb::EPI_ROPE, 0};
    g.rope.out = (hb::bf16*)qkv; g.rope.ldc = N; g.rope.cs = cs; g.rope.slots = slot_mapping;
    g.rope.k_cache = (hb::bf16*)k_cache; g.rope.v_cache = (hb::bf16*)v_cache;
    g.rope.Hq = Hq; g.rope.Hkv = Hkv; g.rope.D = D; g.rope.page_size = page_size;
    e = hb::gemm_bf16_tn(0, g);
  }
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaFree(cs);
  return kret(e);
}
static int sample_impl(const float* logits, int ldl, const float* temperature, const uint64_t* seed, const int32_t* top_k,
                       const float* top_p, int32_t* out, int B, int V) {
  void* scratch = nullptr;
  cudaError_t e = cudaMalloc(&scratch, hb::sample_scratch_bytes(B, V));
  if (e != cudaSuccess) return kret(e);
  e = hb::sample_tokens(0, logits, ldl, temperature, seed, out, B, V, scratch, top_k, top_p);
  cudaError_t e2 = cudaDeviceSynchronize();
  cudaFree(scratch);
  return kret(e != cudaSuccess ? e : e2);
}
int hbk_sample(const float* logits, int ldl, const float* temperature, const uint64_t* seed, int32_t* out, int B, int V) {
  return sample_impl(logits, ldl, temperature, seed, nullptr, nullptr, out, B, V);
}
int hbk_sample_filtered(const float* logits, int ldl, const float* temperature, const uint64_t* seed, const int32_t* top_k,
                        const float* top_p, int32_t* out, int B, int V) {
  return sample_impl(logits, ldl, temperature, seed, top_k, top_p, out, B, V);
}
int hbk_apply_penalties(float* logits, int ldl, const int32_t* pen_off, const void* pen, int B, int V) {
  return kret(hb::apply_penalties(0, logits, ldl, pen_off, pen, B, V));
}
int hbk_logprob_topk(const float* logits, int ldl, int V, const int32_t* sampled, const int32_t* width, int32_t* out_ids,
                     float* out_lp, int B, int max_width) {
  return kret(hb::logprob_topk(0, logits, ldl, V, sampled, width, out_ids, out_lp, B, max_width));
}
int hbk_cls_pool_l2(const void* x, const int32_t* first_row, float* out, int B, int H) {
  return kret(hb::cls_pool_l2(0, (const hb::bf16*)x, first_row, out, B, H));
}
static hb::AttnPrefillArgs mk_prefill(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out,
                                      int ldo, const int32_t* cu, int B, int T, int max_seqlen, int Hq, int Hkv, int D,
                                      int causal, float scale) {
  hb::AttnPrefillArgs a{};
  a.q = (const hb::bf16*)q; a.ldq = ldq;
  a.k = (const hb::bf16*)k; a.ldk = ldk;
  a.v = (const hb::bf16*)v; a.ldv = ldv;
  a.out = (hb::bf16*)out; a.ldo = ldo;
  a.cu_seqlens = cu;
  a.B = B; a.T = T; a.max_seqlen = max_seqlen;
  a.Hq = Hq; a.Hkv = Hkv; a.D = D; a.causal = causal; a.scale = scale;
  return a;
}
int hbk_attn_prefill(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo,
                     const int32_t* cu, int B, int T, int max_seqlen, int Hq, int Hkv, int D, int causal, float scale) {
  return kret(hb::attn_prefill(0, mk_prefill(q, ldq, k, ldk, v, ldv, out, ldo, cu, B, T, max_seqlen, Hq, Hkv, D, causal, scale)));
}
int hbk_attn_prefill_paged(const void* q, int ldq, const void* k_cache, const void* v_cache, const int32_t* page_table,
                           int max_pages, const int32_t* kv_lens, void* out, int ldo, const int32_t* cu, int B, int T,
                           int max_q_len, int Hq, int Hkv, int D, int causal, float scale, int num_pages) {
  hb::AttnPrefillArgs a = mk_prefill(q, ldq, nullptr, 0, nullptr, 0, out, ldo, cu, B, T, max_q_len, Hq, Hkv, D, causal, scale);
  a.k_cache = (const hb::bf16*)k_cache; a.v_cache = (const hb::bf16*)v_cache;
  a.page_table = page_table; a.max_pages = max_pages; a.kv_lens = kv_lens; a.num_pages = num_pages;
  return kret(hb::attn_prefill(0, a));
}
int hbk_attn_naive(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, float* out, int ldo,
                   const int32_t* cu, int B, int T, int max_seqlen, int Hq, int Hkv, int D, int causal, float scale) {
  return kret(hb::attn_naive_check(0, mk_prefill(q, ldq, k, ldk, v, ldv, nullptr, ldo, cu, B, T, max_seqlen, Hq, Hkv, D, causal, scale), out));
}
int hbk_attn_decode(const void* q, int ldq, const void* k_cache, const void* v_cache, const int32_t* page_table,
                    int max_pages, const int32_t* ctx_lens, void* out, int ldo, float* workspace, int B, int Hq, int Hkv,
                    int D, int page_size, int num_splits, float scale, int num_pages) {
  hb::AttnDecodeArgs a{};
  a.num_pages = num_pages;
  a.q = (const hb::bf16*)q; a.ldq = ldq;
  a.k_cache = (const hb::bf16*)k_cache; a.v_cache = (const hb::bf16*)v_cache;
  a.page_table = page_table; a.max_pages = max_pages; a.ctx_lens = ctx_lens;
  a.out = (hb::bf16*)out; a.ldo = ldo; a.workspace = workspace;
  a.B = B; a.Hq = Hq; a.Hkv = Hkv; a.D = D; a.page_size = page_size; a.num_splits = num_splits; a.scale = scale;
  return kret(hb::attn_decode(0, a));
}
size_t hbk_attn_decode_workspace_floats(int B, int Hq, int D, int num_splits) {
  return hb::attn_decode_workspace_floats(B, Hq, D, num_splits);
}

}  // extern "C"
