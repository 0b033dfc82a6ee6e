/* abi_host.c — a host written against include/helix_b200.h and NOTHING else (no Python, no torch, no C++):
 * exactly the calls the Go shim's cgo bindings compile down to (integration/go/b200_runtime.go), in the order
 * Slot.Create / the request handlers make them (api/pkg/runner/slot.go:104-633, openai_chat_handlers.go:100-175).
 * Built by __graft_entry__.build() with gcc; run by tests/test_features_gpu.py, which compares the printed token ids
 * with the Python binding's output for the same deterministic weights.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "helix_b200.h"

#define CK(call)                                                                  \
  do {                                                                            \
    int rc_ = (call);                                                             \
    if (rc_ != HB_OK) {                                                           \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, hb_last_error(eng));          \
      return 1;                                                                   \
    }                                                                             \
  } while (0)

static uint16_t bf16_rne(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += ((u >> 16) & 1u) + 0x7FFFu;
  return (uint16_t)(u >> 16);
}

static hb_engine* eng = NULL;
static int salt = 0;

/* tensor `name` of rows x cols: bf16(((i * 2654435761 + salt * 7919) mod 2001 - 1000) / 25000); vectors are all ones */
static int put(const char* name, size_t rows, size_t cols) {
  const size_t n = rows * cols;
  uint16_t* buf = (uint16_t*)malloc(n * 2);
  if (!buf) return HB_ERR_OOM;
  for (size_t i = 0; i < n; ++i) {
    if (cols == 1) {
      buf[i] = bf16_rne(1.0f);
    } else {
      const uint64_t r = ((uint64_t)i * 2654435761ull + (uint64_t)salt * 7919ull) % 2001ull;
      buf[i] = bf16_rne(((float)r - 1000.0f) / 25000.0f);
    }
  }
  const int rc = hb_model_tensor_set(eng, name, buf, n);
  free(buf);
  ++salt;
  return rc;
}

int main(void) {
  if (hb_abi_version() != HB_ABI_VERSION) {
    fprintf(stderr, "ABI mismatch: header %d, library %d\n", HB_ABI_VERSION, hb_abi_version());
    return 1;
  }
  /* the slot-creation message as the scheduler sends it (types.CreateRunnerSlotRequest) -> engine configuration */
  static const char* create_slot =
      "{\"id\": \"6f1c0e0e-0000-4000-8000-000000000001\", \"attributes\": {\"runtime\": \"vllm\", \"model\": \"tiny/llama\","
      " \"context_length\": 256, \"gpu_index\": 0, \"tensor_parallel_size\": 1,"
      " \"runtime_args\": {\"args\": [\"--max-num-seqs\", \"4\", \"--max-num-batched-tokens\", 256, \"--trust-remote-code\"]}}}";
  hb_engine_cfg cfg;
  hb_slot_info slot;
  if (hb_slot_config(create_slot, 0, &cfg, &slot) != HB_OK || strcmp(slot.model, "tiny/llama") || cfg.max_seqs != 4 ||
      cfg.max_ctx != 256 || cfg.max_batched_tokens != 256 || slot.n_unknown_args != 1) {
    fprintf(stderr, "hb_slot_config: unexpected decode\n");
    return 1;
  }
  if (hb_engine_create(&cfg, &eng) != HB_OK) {
    fprintf(stderr, "hb_engine_create: %s\n", hb_last_error(NULL));
    return 1;
  }
  hb_model_desc d;
  memset(&d, 0, sizeof d);
  d.arch = HB_ARCH_LLAMA;
  d.hidden = 256; d.layers = 2; d.heads = 4; d.kv_heads = 2; d.head_dim = 64; d.ffn = 512; d.vocab = 1000;
  d.max_pos = 4096; d.norm_eps = 1e-5f; d.rope_theta = 500000.0f;
  d.rope_low_freq_factor = 1.0f; d.rope_high_freq_factor = 4.0f;

  uint64_t w = 0, kv = 0, ws = 0;
  CK(hb_memory_estimate(&d, &cfg, &w, &kv, &ws));
  CK(hb_model_load_begin(eng, &d));
  const size_t H = 256, F = 512, V = 1000, QD = 256, KD = 128;
  char name[128];
  CK(put("model.embed_tokens.weight", V, H));
  for (int l = 0; l < d.layers; ++l) {
#define T(suffix, r, c) (snprintf(name, sizeof name, "model.layers.%d." suffix, l), put(name, r, c))
    CK(T("input_layernorm.weight", H, 1));
    CK(T("self_attn.q_proj.weight", QD, H));
    CK(T("self_attn.k_proj.weight", KD, H));
    CK(T("self_attn.v_proj.weight", KD, H));
    CK(T("self_attn.o_proj.weight", H, QD));
    CK(T("post_attention_layernorm.weight", H, 1));
    CK(T("mlp.gate_proj.weight", F, H));
    CK(T("mlp.up_proj.weight", F, H));
    CK(T("mlp.down_proj.weight", H, F));
#undef T
  }
  CK(put("model.norm.weight", H, 1));
  CK(put("lm_head.weight", V, H));
  CK(hb_model_load_finish(eng));
  CK(hb_engine_start(eng));

  /* one greedy chat completion: 32 prompt tokens, 8 generated */
  int32_t prompt[32];
  for (int i = 0; i < 32; ++i) prompt[i] = i + 1;
  hb_sampling sp;
  memset(&sp, 0, sizeof sp);
  sp.max_tokens = 8;
  sp.eos_token = -1;
  sp.top_p = 1.0f;
  sp.logprobs = 3;
  uint64_t req = 0;
  CK(hb_submit(eng, prompt, 32, &sp, &req));
  int32_t out[64];
  int n_total = 0, fin = 0;
  while (!fin) {
    int n = 0;
    hb_wait(eng, req, 10000);
    CK(hb_poll(eng, req, out + n_total, 64 - n_total, &n, &fin));
    n_total += n;
  }
  printf("tokens:");
  for (int i = 0; i < n_total; ++i) printf(" %d", out[i]);
  printf("\n");
  int32_t lp_ids[8 * 3];
  float lp[8 * 3];
  int rows = 0, width = 0;
  CK(hb_logprobs(eng, req, 0, 8, lp_ids, lp, &rows, &width));
  printf("logprobs: rows %d width %d first %d %.4f\n", rows, width, lp_ids[0], lp[0]);
  if (rows != n_total || width != 3 || lp_ids[0] != out[0] || !(lp[0] <= 0.0f)) {
    fprintf(stderr, "log-probability record inconsistent\n");
    return 1;
  }
  CK(hb_release(eng, req));

  /* a cancelled request is retired by the step loop and can then be released (what a client disconnect does) */
  sp.max_tokens = 200;
  sp.logprobs = 0;
  CK(hb_submit(eng, prompt, 32, &sp, &req));
  CK(hb_cancel(eng, req));
  fin = 0;
  while (!fin) {
    int n = 0;
    hb_wait(eng, req, 1000);
    CK(hb_poll(eng, req, out, 64, &n, &fin));
  }
  CK(hb_release(eng, req));

  /* embeddings through the same engine (decoder embedder: last-token pooling + L2) */
  int32_t offs[3] = {0, 5, 32};
  float vec[2 * 256];
  CK(hb_embed(eng, prompt, offs, 2, vec));
  double nn = 0;
  for (int i = 0; i < 256; ++i) nn += (double)vec[i] * vec[i];
  printf("embed_norm: %.4f\n", sqrt(nn));

  unsigned char id[HB_REPLICA_ID_BYTES];
  const int rid_rc = hb_replica_unique_id(id);
  printf("replica_id_rc: %d\n", rid_rc);
  if (rid_rc == HB_OK) {
    double sec = 0;
    CK(hb_model_load_broadcast(eng, &d, id, 0, 1, &sec));
    printf("broadcast_s: %.6f\n", sec);
  }

  hb_stats st;
  CK(hb_get_stats(eng, &st));
  printf("stats: weights %llu est %llu launches %llu free %d/%d cuda_error %d\n", (unsigned long long)st.weights_bytes,
         (unsigned long long)w, (unsigned long long)st.kernel_launches, st.kv_pages_free, st.kv_pages_total, st.cuda_error);
  if (st.weights_bytes != w || st.kv_pages_free != st.kv_pages_total || st.cuda_error || st.kernel_launches == 0) {
    fprintf(stderr, "stats inconsistent\n");
    return 1;
  }
  CK(hb_engine_stop(eng));
  hb_engine_destroy(eng);
  printf("status: ok\n");
  return 0;
}
