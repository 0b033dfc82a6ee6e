// One ModelInstance on one GPU: weight arena, paged KV pool, workspaces, request queues and the
// continuous-batching step loop.  This is the in-process replacement for the backend child process a
// reference Runtime owns (api/pkg/runner/slot.go:46-57); request admission mirrors what the reference
// delegates to vLLM's scheduler (--max-num-seqs / paged KV), page bookkeeping is exact integer work.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <map>
#include <unordered_map>
#include <vector>

#include "../../include/helix_b200.h"
#include "kernels.h"
#include "model.h"

namespace hb {

enum class ReqState { WAITING, RUNNING, FINISHED, CANCELLED, FAILED };

struct Request {
  uint64_t id = 0;
  std::vector<int32_t> prompt;
  std::vector<int32_t> out;
  hb_sampling sp{};
  ReqState state = ReqState::WAITING;
  bool cancel_flag = false;
  std::vector<int32_t> pages;
  int kv_len = 0;       // tokens whose K/V are in the cache
  int prefilled = 0;    // prompt tokens already prefilled (chunked prefill of prompts longer than one step's budget)
  int chunk = 0;        // prompt tokens scheduled in the current prefill step
  int registered = 0;   // leading pages already content-addressed (prefix cache)
  uint64_t chain_key = 0;  // content key of page `registered - 1`
  size_t polled = 0;    // tokens already handed to the caller
  std::vector<float> step_logits;    // rows of [vocab] (HB_CAPTURE_STEP_LOGITS)
  std::vector<float> prompt_logits;  // [n_prompt][vocab] (HB_CAPTURE_PROMPT_LOGITS)
  std::unordered_map<int32_t, int32_t> counts;  // generated token -> occurrences (presence / frequency penalties)
  std::vector<int32_t> lp_ids;       // [generated][sp.logprobs]
  std::vector<float> lp_vals;
  int preempted = 0;    // times this sequence was evicted and re-queued (pages reclaimed, KV recomputed on re-admission)
  int target = 0;       // tokens that must have K/V before the next token is sampled (prompt [+ generated, after a preemption])
  bool in_running = false;
};

struct StepLayout {
  size_t tokens, positions, slots, cu, last, ctx, pt, temp, seed, topk, topp, lpw, pen_off, pen, total;
};
enum StepFlags : int { STEP_FILTER = 1, STEP_PENALTY = 2, STEP_LOGPROBS = 4 };

class Engine {
 public:
  explicit Engine(const hb_engine_cfg& cfg);
  ~Engine();
  int init();  // device, stream

  int load_begin(const hb_model_desc& d);
  int tensor_set(const char* name, const void* host_bf16, size_t n);
  int load_finish();
  int load_random(const hb_model_desc& d, uint64_t seed);
  int load_gguf(const char* path);
  int weights_arena(void** p, size_t* bytes);

  int start();
  int stop();
  int step(int* did_work);
  int submit(const int32_t* toks, int n, const hb_sampling* sp, uint64_t* id);
  int poll(uint64_t id, int32_t* out, int cap, int* n_out, int* finished);
  int wait(uint64_t id, int timeout_ms);
  int cancel(uint64_t id);
  int release(uint64_t id);
  int captured(uint64_t id, int which, float* out, size_t cap, int* rows);
  int embed(const int32_t* toks, const int32_t* offsets, int nseq, float* out);
  int logprobs(uint64_t id, int first_row, int max_rows, int32_t* ids, float* lps, int* rows, int* width);
  static int replica_unique_id(void* id);
  int load_broadcast(const hb_model_desc& d, const void* id, int rank, int world, double* seconds);
  int stats(hb_stats* s);
  int set_profile(bool on);
  int set_mixed(int on, int tokens) {
    std::lock_guard<std::mutex> g(mu_);
    cfg_.decode_with_prefill = on;
    cfg_.mixed_step_tokens = tokens < 0 ? 0 : tokens;
    return HB_OK;
  }
  const char* last_error();

  static void estimate(const hb_model_desc& d, const hb_engine_cfg& c, uint64_t* w, uint64_t* kv, uint64_t* ws);

 private:
  int fail(int code, const std::string& msg);
  int fail_cuda(cudaError_t e, const char* what);
  size_t workspace_bytes() const;
  int alloc_runtime();
  void free_all();
  void loop();
  void finish_request(Request* r, ReqState st);
  void preempt(Request* v);
  void ensure_decode_pages();
  StepLayout layout(int T, int B, size_t pen_entries = 0) const;
  int forward_llama(int T, int B, bool prefill, int max_seqlen, const StepLayout& L, bool all_logits, bool paged = false,
                    float* pool_out = nullptr);
  int forward_bert(int T, int B, int max_seqlen, const StepLayout& L, float* d_out);
  int run_prefill(std::vector<Request*>& batch);
  int run_decode(std::vector<Request*>& batch);
  int decode_splits(int B) const;
  int forward_llama_decode(int B, const StepLayout& L);
  int forward_llama_decode_fused(int B, const StepLayout& L);
  bool fused_decode_ok() const;
  static size_t fused_counter_ints(const hb_model_desc& d);
  size_t skinny_ws_bytes(int sms) const;
  // profiling spans
  struct Span { int cat; cudaEvent_t a, b; double work; };
  cudaEvent_t take_event();
  void span_begin(int cat, double work);
  void span_end();
  int drain_spans();
  bool profile_ = false;
  bool step_is_prefill_ = true;
  double attn_flops_ = 0, attn_bytes_ = 0;  // per-layer attention work of the current step
  std::vector<cudaEvent_t> ev_pool_;
  size_t ev_next_ = 0;
  std::vector<Span> spans_;
  cudaEvent_t fwd_a_ = nullptr, fwd_b_ = nullptr;
  double gpu_ms_prefill_ = 0, gpu_ms_decode_ = 0;
  double prof_ms_[8] = {0}, prof_work_[8] = {0};
  uint64_t prof_launches_[8] = {0};

  hb_engine_cfg cfg_;
  Model model_;
  bool load_open_ = false, loaded_ = false;
  cudaStream_t stream_ = nullptr;
  void* green_ctx_ = nullptr;  // sm_partition: the stream lives in a green context of sm_budget SMs
  int page_ = 64, max_pages_per_seq_ = 0, b_cap_ = 0, t_cap_ = 0;
  uint64_t budget_ = 0;

  // device memory
  bf16* kv_ = nullptr;  // [layers][2][num_pages][Hkv][page][D]
  size_t kv_bytes_ = 0;
  int num_pages_ = 0;
  std::vector<int32_t> free_pages_;
  // ---- prefix cache (mu_ held): full pages are content-addressed by the hash chain of their tokens
  struct PageMeta {
    int ref = 0;            // sequences holding the page
    bool cached = false;    // registered in cache_
    uint64_t key = 0, parent = 0;
    uint64_t tick = 0;      // LRU stamp while unreferenced
    std::vector<int32_t> toks;  // the page's 64 tokens (hits are verified, not trusted to the hash)
  };
  std::vector<PageMeta> pmeta_;
  std::unordered_map<uint64_t, int32_t> cache_;   // content key -> page
  std::map<uint64_t, int32_t> lru_;               // tick -> unreferenced cached page (oldest first)
  uint64_t tick_ = 0, prefix_hit_tokens_ = 0, preemptions_ = 0, steps_mixed_ = 0;
  int32_t take_page();                 // free list first, then evict the oldest unreferenced cached page
  void drop_page(int32_t pg);          // a sequence lets go of a page
  int pages_available() const { return (int)(free_pages_.size() + lru_.size()); }
  void register_full_pages(Request* r);
  int32_t token_at(const Request* r, int pos) const {
    return pos < (int)r->prompt.size() ? r->prompt[pos] : r->out[pos - (int)r->prompt.size()];
  }
  uint8_t* ws_ = nullptr;
  size_t ws_bytes_ = 0;
  bf16 *x_ = nullptr, *xn_ = nullptr, *qkv_ = nullptr, *attn_ = nullptr, *h_ = nullptr;
  float* rope_cs_ = nullptr;  // [t_cap][head_dim] cos | sin of the current step's positions (EPI_ROPE)
  float* logits_ = nullptr;
  float* dec_ws_ = nullptr;
  float* skinny_ws_ = nullptr;  // fp32 partial slabs of the decode-step GEMMs
  SkinnyPlan plan_qkv_{}, plan_o_{}, plan_gu_{}, plan_down_{}, plan_head_{};
  int skinny_max_b_ = 0;
  int32_t* sampled_ = nullptr;
  void* sample_ws_ = nullptr;
  int step_decode_rows_ = 0;    // mixed step: the last N batch entries are decode rows (attention through the decode kernel)
  int step_flags_ = 0;          // StepFlags of the current step: top-k/top-p filter, penalties, log-probabilities
  size_t pen_cap_ = 0;          // penalty entries the step block can hold
  unsigned long long* dec_trace_ = nullptr;  // HB_DEC_TRACE timeline buffer (debug)
  int *cnt_qkv_ = nullptr, *cnt_o_ = nullptr, *cnt_gu_ = nullptr, *cnt_down_ = nullptr, *cnt_head_ = nullptr;  // tile arrival counters
  float* ss_ = nullptr;         // [hidden/128][256] per-tile sums of x^2 (RMSNorm statistics carried between finishers)
  int* dep_ = nullptr;          // [9 * layers + 2] release/acquire dependency flags of the decode chain (kernels.h DepSig)
  int* sig_ = nullptr;          // [5 * layers + 1] HBM hand-over counters of the decode step (kernels.h StreamSig)
  int32_t* lp_ids_ = nullptr;   // [b_cap][HB_MAX_LOGPROBS]
  float* lp_vals_ = nullptr;
  int32_t* h_lp_ids_ = nullptr;
  float* h_lp_vals_ = nullptr;
  size_t fill_sampling(Request* const* batch, int B, const StepLayout& L0, int T);  // per-row sampler inputs; returns pen entries
  int sample_step(int B, const StepLayout& L);  // penalties -> sample -> logprobs on logits_
  int collect_logprobs(Request* const* batch, int B, bool prefill);
  uint8_t* d_step_ = nullptr;  // per-step int/float inputs (layout())
  uint8_t* h_step_ = nullptr;  // pinned mirror
  int32_t* h_sampled_ = nullptr;
  size_t step_bytes_ = 0;
  float* all_logits_ = nullptr;  // lazily allocated [cap_rows][vocab] for HB_CAPTURE_PROMPT_LOGITS
  int all_logits_rows_ = 0;
  float* h_logits_ = nullptr;  // pinned, lazily sized
  size_t h_logits_floats_ = 0;
  float* d_embed_out_ = nullptr;  // [b_cap][hidden] fp32
  float* h_embed_out_ = nullptr;

  // decode CUDA graphs keyed by batch size
  std::unordered_map<int, cudaGraphExec_t> graphs_;
  std::unordered_map<int, uint64_t> graph_kernels_;  // kernels inside each captured step

  // queues
  std::mutex mu_;        // request tables
  std::mutex gpu_mu_;    // serialises GPU steps (step loop vs hb_embed)
  std::condition_variable cv_work_, cv_out_;
  std::deque<Request*> waiting_;
  std::vector<Request*> running_;
  std::unordered_map<uint64_t, std::unique_ptr<Request>> reqs_;
  uint64_t next_id_ = 1;
  std::thread thread_;
  std::atomic<bool> stop_{false}, closing_{false};
  std::atomic<int> waiters_{0};  // threads inside wait(): the destructor wakes them and lets them leave first
  bool thread_running_ = false;

  std::mutex err_mu_;
  std::string last_error_;
  std::atomic<int> cuda_error_{0};
  std::atomic<uint64_t> launches_{0}, graph_launches_{0}, steps_prefill_{0}, steps_decode_{0}, tok_prefill_{0},
      tok_decode_{0};
};

}  // namespace hb
