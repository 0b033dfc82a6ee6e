#include "engine.h"

#include "gguf.h"
#include "launch.h"
#include "nccl_dl.h"
#include "tma_host.h"

#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>

namespace hb {

namespace {

size_t al256(size_t x) { return (x + 255) & ~size_t(255); }
size_t al16(size_t x) { return (x + 15) & ~size_t(15); }

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// uniform in [-a, a] with a = sqrt(3)*std  (std 0.02 like the HF initializer_range)
__global__ void fill_random_kernel(bf16* p, size_t n, uint64_t seed, float amp) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint64_t h = mix64(seed ^ (i * 0xD1B54A32D192ED03ull));
    const float u = (float)(h >> 40) * (1.0f / 16777216.0f) * 2.0f - 1.0f;
    p[i] = __float2bfloat16(u * amp);
  }
}
__global__ void fill_const_kernel(bf16* p, size_t n, float v) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = __float2bfloat16(v);
}

thread_local std::string g_create_error;

}  // namespace

#define CU(expr)                                         \
  do {                                                   \
    cudaError_t e_ = (expr);                             \
    if (e_ != cudaSuccess) return fail_cuda(e_, #expr);  \
  } while (0)
#define LAUNCH(expr)                                     \
  do {                                                   \
    cudaError_t e_ = (expr);                             \
    launches_.fetch_add(1, std::memory_order_relaxed);   \
    if (e_ != cudaSuccess) return fail_cuda(e_, #expr);  \
  } while (0)

Engine::Engine(const hb_engine_cfg& cfg) : cfg_(cfg) {
  if (cfg_.max_seqs <= 0) cfg_.max_seqs = 256;
  if (cfg_.max_ctx <= 0) cfg_.max_ctx = 8192;
  if (cfg_.max_batched_tokens <= 0) cfg_.max_batched_tokens = 16384;
  if (cfg_.kv_page_size <= 0) cfg_.kv_page_size = 64;
  if (cfg_.mixed_step_tokens < 0) cfg_.mixed_step_tokens = 0;
  page_ = cfg_.kv_page_size;
}

Engine::~Engine() {
  stop();
  {
    // Runtime.Stop may race with request handlers still parked in hb_wait (the reference's Stop just kills a child
    // process): fail every open request, wake the waiters and let them leave before anything is freed
    std::lock_guard<std::mutex> g(mu_);
    closing_.store(true);
    for (Request* r : waiting_) finish_request(r, ReqState::CANCELLED);
    waiting_.clear();
    for (Request* r : running_) finish_request(r, ReqState::CANCELLED);
    running_.clear();
    cv_out_.notify_all();
  }
  while (waiters_.load() > 0) {
    cv_out_.notify_all();
    std::this_thread::sleep_for(std::chrono::microseconds(100));
  }
  free_all();
  if (stream_) {
    cudaSetDevice(cfg_.device);
    for (cudaEvent_t e : ev_pool_) cudaEventDestroy(e);
    if (fwd_a_) cudaEventDestroy(fwd_a_);
    if (fwd_b_) cudaEventDestroy(fwd_b_);
    cudaStreamDestroy(stream_);
    if (green_ctx_) destroy_partition(green_ctx_);
  }
}

int Engine::fail(int code, const std::string& msg) {
  std::lock_guard<std::mutex> g(err_mu_);
  last_error_ = msg;
  return code;
}
int Engine::fail_cuda(cudaError_t e, const char* what) {
  cuda_error_.store((int)e);
  return fail(HB_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what + " [" +
                               std::string(tmap_last_error()) + "]");
}
const char* Engine::last_error() {
  // copy under the lock into a per-thread buffer: the returned pointer stays valid for the calling thread even if
  // another thread records a newer error meanwhile
  thread_local std::string tl;
  std::lock_guard<std::mutex> g(err_mu_);
  tl = last_error_;
  return tl.c_str();
}

int Engine::init() {
  if (page_ != 64) return fail(HB_ERR_INVALID, "kv_page_size must be 64");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) return fail(HB_ERR_CUDA, "no CUDA device available (this runtime has no CPU fallback)");
  if (cfg_.device < 0 || cfg_.device >= n) return fail(HB_ERR_INVALID, "device ordinal out of range");
  CU(cudaSetDevice(cfg_.device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, cfg_.device));
  if (prop.major != 10) return fail(HB_ERR_CUDA, "device is not sm_100 (Blackwell B200); kernels are sm_100a-only");
  int dev_sms = 0;
  CU(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, cfg_.device));
  if (cfg_.sm_budget < 0 || cfg_.sm_budget > dev_sms) cfg_.sm_budget = 0;
  if (cfg_.sm_budget > 0 && cfg_.sm_budget < 16) return fail(HB_ERR_INVALID, "sm_budget must be 0 or >= 16 SMs");
  if (cfg_.sm_partition && cfg_.sm_budget > 0) {
    // hardware-enforced share of the GPU for this ModelInstance (CUDA green context): kernels on this stream can only
    // be scheduled on the granted SMs, whatever the other engines on the device launch
    const char* why = nullptr;
    int granted = 0;
    CU(cudaFree(nullptr));  // the primary context must exist before a green context is carved out of it
    if (!create_partition_stream(cfg_.device, cfg_.sm_budget, cfg_.stream_priority, &stream_, &green_ctx_, &granted, &why))
      return fail(HB_ERR_INVALID, std::string("sm_partition: ") + (why ? why : "unavailable"));
    cfg_.sm_budget = granted;  // the driver rounds the request up to its SM granularity
  } else {
    int lo = 0, hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CU(cudaStreamCreateWithPriority(&stream_, cudaStreamNonBlocking, cfg_.stream_priority > 0 ? hi : lo));
  }
  CU(cudaEventCreate(&fwd_a_));
  CU(cudaEventCreate(&fwd_b_));
  CU(kernels_init());  // max-dynamic-smem attributes for every instantiation (never inside a graph capture)
  return HB_OK;
}

void Engine::free_all() {
  if (cfg_.device >= 0) cudaSetDevice(cfg_.device);
  for (auto& kv : graphs_) cudaGraphExecDestroy(kv.second);
  graphs_.clear();
  auto fr = [](auto*& p) {
    if (p) cudaFree(p);
    p = nullptr;
  };
  fr(model_.arena);
  fr(model_.inv_freq);
  fr(kv_);
  fr(ws_);
  fr(d_step_);
  fr(all_logits_);
  fr(d_embed_out_);
  fr(dec_trace_);
  auto frh = [](auto*& p) {
    if (p) cudaFreeHost(p);
    p = nullptr;
  };
  frh(h_step_);
  frh(h_sampled_);
  frh(h_lp_ids_);
  frh(h_lp_vals_);
  frh(h_logits_);
  frh(h_embed_out_);
  loaded_ = false;
  load_open_ = false;
}

// ------------------------------------------------------------------ sizing
size_t Engine::workspace_bytes() const {
  const hb_model_desc& d = model_.d;
  const size_t T = t_cap_, H = d.hidden;
  size_t b = 0;
  b += 2 * al256(T * H * 2);                                  // x, xn
  b += al256(T * (size_t)model_.qkv_cols() * 2);              // qkv
  b += al256(T * (size_t)d.heads * d.head_dim * 2);           // attn
  b += al256(T * (size_t)d.ffn * 2);                          // h
  if (d.arch == HB_ARCH_LLAMA) {
    b += al256((size_t)cfg_.max_seqs * d.vocab * 4);          // logits
    b += al256(attn_decode_workspace_floats(cfg_.max_seqs, d.heads, d.head_dim, 16) * 4);
    b += al256((size_t)cfg_.max_seqs * 4);                    // sampled
    b += al256(sample_scratch_bytes(cfg_.max_seqs, d.vocab)); // argmax partials
    b += 2 * al256((size_t)cfg_.max_seqs * HB_MAX_LOGPROBS * 4);  // log-probability records of one step
    b += al256((size_t)(5 * d.layers + 1 + 9 * d.layers + 2) * sizeof(int));  // hand-over counters + dependency flags of the decode step
    b += al256(fused_counter_ints(d) * sizeof(int));              // tile arrival counters of the fused decode GEMMs
    b += al256((size_t)((d.hidden + 127) / 128) * kSkinnySsStride * sizeof(float));  // RMSNorm per-tile partials
    b += al256(skinny_ws_bytes(148));                         // decode GEMM partial slabs
    b += al256(T * (size_t)d.head_dim * 4);                   // cos/sin table of a step (QKV epilogue)
  }
  return b;
}

size_t Engine::fused_counter_ints(const hb_model_desc& d) {
  const size_t qkv = (size_t)(d.heads + 2 * d.kv_heads) * d.head_dim;
  return (qkv + 127) / 128 + 2 * ((size_t)(d.hidden + 127) / 128) + ((size_t)2 * d.ffn + 127) / 128 + ((size_t)d.vocab + 127) / 128;
}

size_t Engine::skinny_ws_bytes(int sms) const {
  const hb_model_desc& d = model_.d;
  const size_t B = std::min(cfg_.max_seqs, 256);
  const int H = d.hidden, QD = d.heads * d.head_dim, QKV = (d.heads + 2 * d.kv_heads) * d.head_dim, F = d.ffn;
  size_t f = 0;
  f = std::max(f, (size_t)gemm_skinny_max_segs(QKV, H, sms) * B * QKV);
  f = std::max(f, (size_t)gemm_skinny_max_segs(H, QD, sms) * B * H);
  f = std::max(f, (size_t)gemm_skinny_max_segs(2 * F, H, sms) * B * 2 * F);
  f = std::max(f, (size_t)gemm_skinny_max_segs(H, F, sms) * B * H);
  f = std::max(f, (size_t)gemm_skinny_max_segs(d.vocab, H, sms) * B * d.vocab);
  return f * 4;
}

void Engine::estimate(const hb_model_desc& d, const hb_engine_cfg& c_in, uint64_t* w, uint64_t* kv, uint64_t* ws) {
  Engine tmp(c_in);
  tmp.model_.d = d;
  tmp.t_cap_ = tmp.cfg_.max_batched_tokens;
  if (d.arch == HB_ARCH_BERT) tmp.t_cap_ = std::max(tmp.t_cap_, std::min(tmp.cfg_.max_ctx, d.max_pos));
  if (w) *w = arena_bytes_for(d);
  if (ws) *ws = tmp.workspace_bytes();
  if (kv) {
    if (d.arch == HB_ARCH_LLAMA) {
      const uint64_t pages_per_seq = (tmp.cfg_.max_ctx + 63) / 64;
      const uint64_t page_bytes = (uint64_t)d.layers * 2 * d.kv_heads * 64 * d.head_dim * 2;
      *kv = (uint64_t)tmp.cfg_.max_seqs * pages_per_seq * page_bytes;
    } else {
      *kv = 0;
    }
  }
}

StepLayout Engine::layout(int T, int B, size_t pen_entries) const {
  StepLayout L;
  size_t o = 0;
  L.tokens = o; o = al16(o + 4 * (size_t)T);
  L.positions = o; o = al16(o + 4 * (size_t)T);
  L.slots = o; o = al16(o + 4 * (size_t)T);
  L.cu = o; o = al16(o + 4 * (size_t)(B + 1));
  L.last = o; o = al16(o + 4 * (size_t)B);
  L.ctx = o; o = al16(o + 4 * (size_t)B);
  L.pt = o; o = al16(o + 4 * (size_t)B * max_pages_per_seq_);
  L.temp = o; o = al16(o + 4 * (size_t)B);
  L.seed = o; o = al16(o + 8 * (size_t)B);
  L.topk = o; o = al16(o + 4 * (size_t)B);
  L.topp = o; o = al16(o + 4 * (size_t)B);
  L.lpw = o; o = al16(o + 4 * (size_t)B);
  L.pen_off = o; o = al16(o + 4 * (size_t)(B + 1));
  L.pen = o; o = al16(o + 8 * pen_entries);  // {int32 token; float value} entries, last so no offset depends on their number
  L.total = o;
  return L;
}

// ------------------------------------------------------------------ model load
int Engine::load_begin(const hb_model_desc& d) {
  std::lock_guard<std::mutex> g(gpu_mu_);
  if (loaded_ || load_open_) return fail(HB_ERR_STATE, "model already loaded on this engine");
  const std::string why = validate_desc(d);
  if (!why.empty()) return fail(HB_ERR_INVALID, "unsupported model description: " + why);
  CU(cudaSetDevice(cfg_.device));
  size_t free_b = 0, total_b = 0;
  CU(cudaMemGetInfo(&free_b, &total_b));
  budget_ = cfg_.memory_budget_bytes ? cfg_.memory_budget_bytes : (free_b > (1ull << 30) ? free_b - (1ull << 30) : free_b);
  model_.d = d;
  model_.arena_bytes = arena_bytes_for(d);
  if (model_.arena_bytes > budget_) return fail(HB_ERR_OOM, "weights alone exceed the memory budget");
  if (model_.arena_bytes > free_b) return fail(HB_ERR_OOM, "weights exceed free device memory");
  CU(cudaMalloc(&model_.arena, model_.arena_bytes));
  layout_model(model_);
  load_open_ = true;
  return HB_OK;
}

int Engine::tensor_set(const char* name, const void* host, size_t n) {
  std::lock_guard<std::mutex> g(gpu_mu_);
  if (!load_open_) return fail(HB_ERR_STATE, "hb_model_tensor_set outside load_begin/finish");
  auto it = model_.placements.find(name);
  if (it == model_.placements.end()) return fail(HB_ERR_NOT_FOUND, std::string("unknown tensor name: ") + name);
  const Placement& p = it->second;
  if (n != p.rows * p.cols) return fail(HB_ERR_INVALID, std::string("tensor size mismatch for ") + name);
  CU(cudaSetDevice(cfg_.device));
  if (p.mode == 0) {
    CU(cudaMemcpy(p.dst, host, n * 2, cudaMemcpyHostToDevice));
  } else {
    // gate (mode 1) / up (mode 2) rows interleaved per 128-row block: block t -> rows [t*256 + (mode-1)*128, +128)
    const size_t blocks = p.rows / 128;
    const bf16* src = static_cast<const bf16*>(host);
    for (size_t t = 0; t < blocks; ++t) {
      bf16* dst = p.dst + (t * 256 + (p.mode - 1) * 128) * p.cols;
      CU(cudaMemcpyAsync(dst, src + t * 128 * p.cols, 128 * p.cols * 2, cudaMemcpyHostToDevice, stream_));
    }
    CU(cudaStreamSynchronize(stream_));
  }
  model_.filled[name] = true;
  return HB_OK;
}

int Engine::alloc_runtime() {
  const hb_model_desc& d = model_.d;
  SmLimitScope sm_scope(cfg_.sm_budget);  // the stream-K plans of the decode GEMMs are cut for this engine's SM share
  t_cap_ = cfg_.max_batched_tokens;
  if (d.arch == HB_ARCH_BERT && cfg_.max_ctx > d.max_pos) cfg_.max_ctx = d.max_pos;
  if (d.arch == HB_ARCH_BERT && t_cap_ < cfg_.max_ctx) t_cap_ = cfg_.max_ctx;  // an encoder sequence is never split
  b_cap_ = std::max(cfg_.max_seqs, d.arch == HB_ARCH_BERT ? 4096 : cfg_.max_seqs);
  max_pages_per_seq_ = (cfg_.max_ctx + page_ - 1) / page_;

  ws_bytes_ = workspace_bytes();
  CU(cudaMalloc(&ws_, ws_bytes_));
  uint8_t* p = ws_;
  auto take = [&](size_t bytes) {
    uint8_t* r = p;
    p += al256(bytes);
    return r;
  };
  const size_t T = t_cap_, H = d.hidden;
  x_ = (bf16*)take(T * H * 2);
  xn_ = (bf16*)take(T * H * 2);
  qkv_ = (bf16*)take(T * (size_t)model_.qkv_cols() * 2);
  attn_ = (bf16*)take(T * (size_t)d.heads * d.head_dim * 2);
  h_ = (bf16*)take(T * (size_t)d.ffn * 2);
  if (d.arch == HB_ARCH_LLAMA) {
    logits_ = (float*)take((size_t)cfg_.max_seqs * d.vocab * 4);
    dec_ws_ = (float*)take(attn_decode_workspace_floats(cfg_.max_seqs, d.heads, d.head_dim, 16) * 4);
    sampled_ = (int32_t*)take((size_t)cfg_.max_seqs * 4);
    sample_ws_ = take(sample_scratch_bytes(cfg_.max_seqs, d.vocab));
    sig_ = (int*)take((size_t)(5 * d.layers + 1 + 9 * d.layers + 2) * sizeof(int));
    dep_ = sig_ + (5 * d.layers + 1);  // dependency flags of the decode chain follow the hand-over counters (one memset)
    {
      int* cnt = (int*)take(fused_counter_ints(d) * sizeof(int));
      CU(cudaMemset(cnt, 0, fused_counter_ints(d) * sizeof(int)));
      const int tq = (model_.qkv_cols() + 127) / 128, th = (d.hidden + 127) / 128, tg = (2 * d.ffn + 127) / 128;
      cnt_qkv_ = cnt;
      cnt_o_ = cnt_qkv_ + tq;
      cnt_gu_ = cnt_o_ + th;
      cnt_down_ = cnt_gu_ + tg;
      cnt_head_ = cnt_down_ + th;
      ss_ = (float*)take((size_t)th * kSkinnySsStride * sizeof(float));
    }
    if (getenv("HB_DEC_TRACE")) {  // debug timeline: 16 %globaltimer stamps per streaming kernel (tools/dec_trace.py)
      CU(cudaMalloc(&dec_trace_, (size_t)(5 * d.layers + 1) * 16 * 8));
      CU(cudaMemset(dec_trace_, 0, (size_t)(5 * d.layers + 1) * 16 * 8));
    }
    lp_ids_ = (int32_t*)take((size_t)cfg_.max_seqs * HB_MAX_LOGPROBS * 4);
    lp_vals_ = (float*)take((size_t)cfg_.max_seqs * HB_MAX_LOGPROBS * 4);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg_.device);
    skinny_ws_ = (float*)take(skinny_ws_bytes(148));
    rope_cs_ = (float*)take(T * (size_t)d.head_dim * 4);
    if (skinny_ws_bytes(sms) > skinny_ws_bytes(148)) return fail(HB_ERR_INVALID, "unexpected SM count for the decode workspace");
    skinny_max_b_ = std::min(cfg_.max_seqs, 256);
    const int QD = d.heads * d.head_dim, QKV = model_.qkv_cols();
    CU(gemm_skinny_plan(QKV, d.hidden, &plan_qkv_));
    CU(gemm_skinny_plan(d.hidden, QD, &plan_o_));
    CU(gemm_skinny_plan(2 * d.ffn, d.hidden, &plan_gu_));
    CU(gemm_skinny_plan(d.hidden, d.ffn, &plan_down_));
    CU(gemm_skinny_plan(d.vocab, d.hidden, &plan_head_));
    {
      const size_t Bm = skinny_max_b_, have = skinny_ws_bytes(148) / 4;
      const size_t need = std::max({gemm_skinny_ws_floats(plan_qkv_, Bm, QKV), gemm_skinny_ws_floats(plan_o_, Bm, d.hidden),
                                    gemm_skinny_ws_floats(plan_gu_, Bm, 2 * d.ffn), gemm_skinny_ws_floats(plan_down_, Bm, d.hidden),
                                    gemm_skinny_ws_floats(plan_head_, Bm, d.vocab)});
      if (need > have) return fail(HB_ERR_INVALID, "decode GEMM plan needs more slab workspace than was sized (sm_budget?)");
    }
  }
  // penalty entries: one per distinct generated token of every running sequence, bounded at 4 Mi (32 MB)
  pen_cap_ = d.arch == HB_ARCH_LLAMA ? std::min<size_t>((size_t)cfg_.max_seqs * cfg_.max_ctx, size_t(4) << 20) : 0;
  step_bytes_ = layout(t_cap_, b_cap_, pen_cap_).total;
  CU(cudaMalloc(&d_step_, step_bytes_));
  CU(cudaMallocHost(&h_step_, step_bytes_));
  CU(cudaMallocHost(&h_sampled_, (size_t)b_cap_ * 4));
  if (d.arch == HB_ARCH_LLAMA) {
    CU(cudaMallocHost(&h_lp_ids_, (size_t)cfg_.max_seqs * HB_MAX_LOGPROBS * 4));
    CU(cudaMallocHost(&h_lp_vals_, (size_t)cfg_.max_seqs * HB_MAX_LOGPROBS * 4));
  }

  if (d.arch == HB_ARCH_LLAMA) {
    const std::vector<float> f = rope_inv_freq(d);
    CU(cudaMalloc(&model_.inv_freq, f.size() * 4));
    CU(cudaMemcpy(model_.inv_freq, f.data(), f.size() * 4, cudaMemcpyHostToDevice));
    // KV pool: whatever the budget leaves, capped at max_seqs full-length sequences
    const uint64_t page_bytes = (uint64_t)d.layers * 2 * d.kv_heads * page_ * d.head_dim * 2;
    const uint64_t used = model_.arena_bytes + ws_bytes_ + step_bytes_;
    if (used >= budget_) return fail(HB_ERR_OOM, "weights + workspace exceed the memory budget");
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    uint64_t avail = std::min<uint64_t>(budget_ - used, free_b > (256ull << 20) ? free_b - (256ull << 20) : 0);
    uint64_t pages = avail / page_bytes;
    const uint64_t want = (uint64_t)cfg_.max_seqs * max_pages_per_seq_;
    if (pages > want) pages = want;
    if (pages < (uint64_t)max_pages_per_seq_)
      return fail(HB_ERR_OOM, "memory budget leaves no room for one full-context sequence of KV cache");
    num_pages_ = (int)pages;
    kv_bytes_ = pages * page_bytes;
    CU(cudaMalloc(&kv_, kv_bytes_));
    // masked positions of a partly-filled page are still multiplied (by an exact 0) in P·V: the pool must never hold
    // NaN/Inf bit patterns left behind by a previous owner of the memory
    CU(cudaMemsetAsync(kv_, 0, kv_bytes_, stream_));
    CU(cudaStreamSynchronize(stream_));
    pmeta_.assign(num_pages_, PageMeta{});
    cache_.clear();
    lru_.clear();
    free_pages_.resize(num_pages_);
    for (int i = 0; i < num_pages_; ++i) free_pages_[i] = num_pages_ - 1 - i;  // pop_back hands out page 0 first
  }
  CU(cudaMalloc(&d_embed_out_, (size_t)b_cap_ * H * 4));  // encoders; decoders serving --task embed (last-token pooling)
  CU(cudaMallocHost(&h_embed_out_, (size_t)b_cap_ * H * 4));
  return HB_OK;
}

int Engine::load_finish() {
  std::lock_guard<std::mutex> g(gpu_mu_);
  if (!load_open_) return fail(HB_ERR_STATE, "hb_model_load_finish without load_begin");
  for (auto& kv : model_.placements)
    if (!model_.filled.count(kv.first)) return fail(HB_ERR_STATE, "tensor never uploaded: " + kv.first);
  CU(cudaSetDevice(cfg_.device));
  int rc = alloc_runtime();
  if (rc != HB_OK) return rc;
  load_open_ = false;
  loaded_ = true;
  return HB_OK;
}

int Engine::load_random(const hb_model_desc& d, uint64_t seed) {
  int rc = load_begin(d);
  if (rc != HB_OK) return rc;
  std::lock_guard<std::mutex> g(gpu_mu_);
  CU(cudaSetDevice(cfg_.device));
  const size_t n = model_.arena_bytes / 2;
  fill_random_kernel<<<148 * 8, 256, 0, stream_>>>(model_.arena, n, seed, 0.02f * 1.7320508f);
  CU(cudaGetLastError());
  for (auto& kv : model_.placements) {
    const Placement& p = kv.second;
    if (p.is_norm_gain) {
      fill_const_kernel<<<8, 256, 0, stream_>>>(p.dst, p.rows * p.cols, 1.0f);
      CU(cudaGetLastError());
    }
    model_.filled[kv.first] = true;
  }
  CU(cudaStreamSynchronize(stream_));
  int rc2 = alloc_runtime();
  if (rc2 != HB_OK) return rc2;
  load_open_ = false;
  loaded_ = true;
  return HB_OK;
}

// GGUF checkpoint -> arena: metadata gives the description, every tensor is dequantised to bf16 on the host and uploaded
// under its HF name (gguf.cpp); q/k projections of llama-architecture files are un-permuted to the rotate-half row order.
int Engine::load_gguf(const char* path) {
  GgufFile g;
  std::string err;
  if (!path || !g.open(path, &err)) return fail(HB_ERR_INVALID, "gguf: " + (path ? err : std::string("null path")));
  hb_model_desc d;
  if (!g.describe(&d, &err)) return fail(HB_ERR_INVALID, "gguf: " + err);
  int rc = load_begin(d);
  if (rc != HB_OK) return rc;
  const bool permute = g.str("general.architecture") == "llama";
  std::vector<float> f;
  std::vector<uint16_t> bits;
  for (const auto& kv : g.tensors()) {
    const std::string hf = gguf_to_hf_name(kv.first);
    if (hf.empty()) continue;  // tensors the engine has no use for
    size_t rows = 0, cols = 0;
    if (!g.read_f32(kv.first, &f, &rows, &cols)) return fail(HB_ERR_INVALID, "gguf: cannot read " + kv.first);
    if (permute && hf.find("self_attn.q_proj.weight") != std::string::npos) gguf_unpermute_rows(f, rows, cols, d.heads);
    if (permute && hf.find("self_attn.k_proj.weight") != std::string::npos) gguf_unpermute_rows(f, rows, cols, d.kv_heads);
    gguf_to_bf16(f, &bits);
    rc = tensor_set(hf.c_str(), bits.data(), bits.size());
    if (rc != HB_OK) return rc;
  }
  return load_finish();
}

int Engine::weights_arena(void** p, size_t* bytes) {
  if (!model_.arena) return fail(HB_ERR_STATE, "no model arena");
  if (p) *p = model_.arena;
  if (bytes) *bytes = model_.arena_bytes;
  return HB_OK;
}

// ------------------------------------------------------------------ replicas: one NCCL broadcast of the arena
int Engine::replica_unique_id(void* id) {
  const char* why = nullptr;
  const NcclApi* n = nccl_api(&why);
  if (!n || !id) return HB_ERR_STATE;
  ncclUniqueId u;
  if (n->GetUniqueId(&u) != ncclSuccess) return HB_ERR_CUDA;
  static_assert(sizeof(ncclUniqueId) == HB_REPLICA_ID_BYTES, "NCCL unique id size");
  memcpy(id, &u, sizeof u);
  return HB_OK;
}

int Engine::load_broadcast(const hb_model_desc& d, const void* id, int rank, int world, double* seconds) {
  if (!id || world < 1 || rank < 0 || rank >= world) return fail(HB_ERR_INVALID, "bad replica rank / world");
  const char* why = nullptr;
  const NcclApi* n = nccl_api(&why);
  if (!n) return fail(HB_ERR_STATE, std::string("NCCL unavailable: ") + (why ? why : ""));
  if (rank == 0) {
    if (!loaded_) return fail(HB_ERR_STATE, "replica root must hold a loaded model before the broadcast");
    if (memcmp(&model_.d, &d, offsetof(hb_model_desc, reserved)) != 0)
      return fail(HB_ERR_INVALID, "replica root: description differs from the loaded model");
  } else {
    int rc = load_begin(d);  // allocates and lays out the arena (fails if a model is already loaded)
    if (rc != HB_OK) return rc;
  }
  std::lock_guard<std::mutex> g(gpu_mu_);
  CU(cudaSetDevice(cfg_.device));
  ncclUniqueId u;
  memcpy(&u, id, sizeof u);
  ncclComm_t comm = nullptr;
  ncclResult_t r = n->CommInitRank(&comm, world, u, rank);
  if (r != ncclSuccess) return fail(HB_ERR_CUDA, std::string("ncclCommInitRank: ") + n->GetErrorString(r));
  // a 1 MB warm-up broadcast of the arena head keeps channel setup out of the measured copy (contents are rewritten below)
  const size_t warm = std::min<size_t>(model_.arena_bytes, size_t(1) << 20);
  r = n->Broadcast(model_.arena, model_.arena, warm, ncclUint8, 0, comm, stream_);
  cudaEvent_t a = nullptr, b = nullptr;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaStreamSynchronize(stream_);
  const auto t0 = std::chrono::steady_clock::now();
  cudaEventRecord(a, stream_);
  if (r == ncclSuccess) r = n->Broadcast(model_.arena, model_.arena, model_.arena_bytes, ncclUint8, 0, comm, stream_);
  cudaEventRecord(b, stream_);
  cudaError_t ce = cudaStreamSynchronize(stream_);
  const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  n->CommDestroy(comm);
  if (r != ncclSuccess) return fail(HB_ERR_CUDA, std::string("ncclBroadcast: ") + n->GetErrorString(r));
  if (ce != cudaSuccess) return fail_cuda(ce, "ncclBroadcast (stream sync)");
  if (seconds) *seconds = std::max((double)ms * 1e-3, 0.0) > 0 ? (double)ms * 1e-3 : wall;
  if (rank != 0) {
    for (auto& kv : model_.placements) model_.filled[kv.first] = true;
    int rc = alloc_runtime();
    if (rc != HB_OK) return rc;
    load_open_ = false;
    loaded_ = true;
  }
  return HB_OK;
}

// ------------------------------------------------------------------ measurement aids
cudaEvent_t Engine::take_event() {
  if (ev_next_ == ev_pool_.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    ev_pool_.push_back(e);
  }
  return ev_pool_[ev_next_++];
}
void Engine::span_begin(int cat, double work) {
  if (!profile_) return;
  Span s{cat, take_event(), take_event(), work};
  cudaEventRecord(s.a, stream_);
  spans_.push_back(s);
}
void Engine::span_end() {
  if (!profile_) return;
  cudaEventRecord(spans_.back().b, stream_);
}
int Engine::drain_spans() {  // stream must be synchronised
  for (const Span& s : spans_) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) {
      prof_ms_[s.cat] += ms;
      prof_work_[s.cat] += s.work;
      prof_launches_[s.cat] += 1;
    }
  }
  spans_.clear();
  ev_next_ = 0;
  return HB_OK;
}
int Engine::set_profile(bool on) {
  std::lock_guard<std::mutex> g(gpu_mu_);
  profile_ = on;
  return HB_OK;
}
#define SPAN(cat, work, expr) \
  do {                        \
    span_begin(cat, work);    \
    LAUNCH(expr);             \
    span_end();               \
  } while (0)

// ------------------------------------------------------------------ forward passes
int Engine::decode_splits(int B) const {
  // two CTAs per SM (106 KB ring each): the fewest splits that put a CTA on ~80 % of the 296 slots.  Measured at
  // batch 32 x 2k context: 1 / 2 / 4 / 8 splits = 6718 / 6673 / 6669 / 6667 tok/s — flat once the machine is covered,
  // so extra splits only add combine work; small batches need them to reach all SMs.
  static const int forced = [] { const char* e = getenv("HB_DECODE_SPLITS"); return e ? atoi(e) : 0; }();  // A/B knob
  if (forced > 0) return std::min(16, forced);
  const int ctas = B * model_.d.kv_heads;
  const int want = (int)(0.8 * 2 * (cfg_.sm_budget > 0 ? cfg_.sm_budget : 148));
  return std::max(1, std::min(16, (want + ctas - 1) / ctas));
}

int Engine::forward_llama(int T, int B, bool prefill, int max_seqlen, const StepLayout& L, bool all_logits, bool paged,
                          float* pool_out) {
  if (!prefill && !all_logits && T == B && B <= skinny_max_b_) return forward_llama_decode(B, L);
  const hb_model_desc& d = model_.d;
  const int H = d.hidden, D = d.head_dim, QD = d.heads * D, KD = d.kv_heads * D, QKV = model_.qkv_cols(), F = d.ffn;
  const int32_t* tokens = (const int32_t*)(d_step_ + L.tokens);
  const int32_t* positions = (const int32_t*)(d_step_ + L.positions);
  const int32_t* slots = (const int32_t*)(d_step_ + L.slots);
  const int32_t* cu = (const int32_t*)(d_step_ + L.cu);
  const int32_t* last = (const int32_t*)(d_step_ + L.last);
  const int32_t* ctx = (const int32_t*)(d_step_ + L.ctx);
  const int32_t* pt = (const int32_t*)(d_step_ + L.pt);
  const size_t layer_kv = (size_t)num_pages_ * d.kv_heads * page_ * D;  // elements per K (or V) plane

  const int gcat = prefill ? 0 : 4;  // GEMM family: FLOPs in prefill steps, weight bytes in decode steps
  auto gwork = [&](double M, double N, double K) { return prefill ? 2.0 * M * N * K : 2.0 * (N * K + M * K + M * N); };
  const double row_bytes = 2.0 * T * H * 2;
  SPAN(3, 2.0 * T * H, embed_gather(stream_, tokens, model_.embed, x_, T, H));
  // RoPE + KV scatter ride in the QKV projection's epilogue (EPI_ROPE) when its 256-column tiles hold whole heads; the
  // step's cos/sin table is built once here and read by every layer.  HB_PREFILL_FUSE_ROPE=0 keeps the separate row kernel.
  static const bool fuse_env = [] { const char* e = getenv("HB_PREFILL_FUSE_ROPE"); return !(e && atoi(e) == 0); }();
  const bool fuse_rope = fuse_env && QKV % 256 == 0 && (D == 64 || D == 128);
  if (fuse_rope) SPAN(3, 1.0 * T * D * 4, rope_table(stream_, positions, model_.inv_freq, rope_cs_, T, D));
  for (int l = 0; l < d.layers; ++l) {
    const LlamaLayerW& w = model_.ll[l];
    bf16* kc = kv_ + (size_t)l * 2 * layer_kv;
    bf16* vc = kc + layer_kv;
    SPAN(3, row_bytes, rmsnorm(stream_, x_, w.attn_norm, xn_, nullptr, T, H, d.norm_eps));
    {
      GemmArgs g{xn_, H, w.wqkv, H, qkv_, QKV, nullptr, 0, w.bqkv, T, QKV, H, w.bqkv ? EPI_BIAS : EPI_NONE, 0};
      if (fuse_rope) {
        g.epi = EPI_ROPE;
        g.rope.out = qkv_; g.rope.ldc = QKV; g.rope.cs = rope_cs_; g.rope.slots = slots;
        g.rope.k_cache = kc; g.rope.v_cache = vc;
        g.rope.Hq = d.heads; g.rope.Hkv = d.kv_heads; g.rope.D = D; g.rope.page_size = page_;
      }
      SPAN(gcat, gwork(T, QKV, H), gemm_bf16_tn(stream_, g));
    }
    if (!fuse_rope)
      SPAN(3, 2.0 * T * (QD + 2.0 * KD) * 2, rope_kv_write(stream_, qkv_, positions, slots, model_.inv_freq, kc, vc, T, d.heads, d.kv_heads, D, page_));
    const int Bd = prefill ? step_decode_rows_ : 0, Bpf = B - Bd, Tpf = T - Bd;  // mixed step: decode rows trail the batch
    if (prefill && Bpf > 0) {
      AttnPrefillArgs a{};
      a.q = qkv_; a.ldq = QKV;
      a.k = qkv_ + QD; a.ldk = QKV;
      a.v = qkv_ + QD + KD; a.ldv = QKV;
      a.out = attn_; a.ldo = QD;
      a.cu_seqlens = cu;
      a.B = Bpf; a.T = Tpf; a.max_seqlen = max_seqlen;
      a.Hq = d.heads; a.Hkv = d.kv_heads; a.D = D;
      a.causal = 1;
      a.scale = 1.0f / sqrtf((float)D);
      if (paged) {  // some sequence continues a partly prefilled prompt: K/V (this chunk's were just written) from the pool
        a.k_cache = kc; a.v_cache = vc;
        a.page_table = pt; a.max_pages = max_pages_per_seq_;
        a.kv_lens = ctx;
        a.num_pages = num_pages_; a.page_size = page_;
      }
      SPAN(1, attn_flops_, attn_prefill(stream_, a));
    }
    if (!prefill || Bd > 0) {
      // decode rows (a whole decode step on this generic path, or the running sequences riding in a mixed step): one query
      // token per sequence over its paged context — the HBM-bound decode-attention kernel, not a 128-row prefill q tile each
      const int b0 = prefill ? Bpf : 0, t0 = prefill ? Tpf : 0, nb = prefill ? Bd : B;
      AttnDecodeArgs a{};
      a.q = qkv_ + (size_t)t0 * QKV; a.ldq = QKV;
      a.k_cache = kc; a.v_cache = vc;
      a.page_table = pt + (size_t)b0 * max_pages_per_seq_; a.max_pages = max_pages_per_seq_;
      a.ctx_lens = ctx + b0;
      a.out = attn_ + (size_t)t0 * QD; a.ldo = QD;
      a.workspace = dec_ws_;
      a.B = nb; a.Hq = d.heads; a.Hkv = d.kv_heads; a.D = D; a.page_size = page_;
      a.num_splits = decode_splits(nb);
      a.scale = 1.0f / sqrtf((float)D);
      a.num_pages = num_pages_;
      SPAN(2, attn_bytes_, attn_decode(stream_, a));
      if (a.num_splits > 1) launches_.fetch_add(1, std::memory_order_relaxed);  // combine kernel
    }
    {
      GemmArgs g{attn_, QD, w.wo, QD, x_, H, x_, H, nullptr, T, H, QD, EPI_RESID, 0};
      SPAN(gcat, gwork(T, H, QD), gemm_bf16_tn(stream_, g));
    }
    SPAN(3, row_bytes, rmsnorm(stream_, x_, w.mlp_norm, xn_, nullptr, T, H, d.norm_eps));
    {
      GemmArgs g{xn_, H, w.wgu, H, h_, F, nullptr, 0, nullptr, T, 2 * F, H, EPI_SWIGLU, 256};
      SPAN(gcat, gwork(T, 2.0 * F, H), gemm_bf16_tn(stream_, g));
    }
    {
      GemmArgs g{h_, F, w.wdown, F, x_, H, x_, H, nullptr, T, H, F, EPI_RESID, 0};
      SPAN(gcat, gwork(T, H, F), gemm_bf16_tn(stream_, g));
    }
  }
  if (pool_out) {  // decoder embedder (--task embed): final-norm hidden state of each sequence's LAST token, L2-normalised
    LAUNCH(rmsnorm(stream_, x_, model_.final_norm, h_, last, B, H, d.norm_eps));
    LAUNCH(cls_pool_l2(stream_, h_, ctx, pool_out, B, H));  // ctx carries 0..B-1 here (rows of h_)
    return HB_OK;
  }
  if (all_logits) {
    LAUNCH(rmsnorm(stream_, x_, model_.final_norm, xn_, nullptr, T, H, d.norm_eps));
    GemmArgs g{xn_, H, model_.lm_head, H, all_logits_, d.vocab, nullptr, 0, nullptr, T, d.vocab, H, EPI_F32, 0};
    LAUNCH(gemm_bf16_tn(stream_, g));
  }
  // last position of every sequence -> final norm -> LM head -> sample
  LAUNCH(rmsnorm(stream_, x_, model_.final_norm, h_, last, B, H, d.norm_eps));
  {
    GemmArgs g{h_, H, model_.lm_head, H, logits_, d.vocab, nullptr, 0, nullptr, B, d.vocab, H, EPI_F32, 0};
    SPAN(gcat, gwork(B, d.vocab, H), gemm_bf16_tn(stream_, g));
  }
  return sample_step(B, L);
}

// penalties -> (top-k / top-p threshold) -> Gumbel-max / argmax -> log-probabilities, all on logits_[B, vocab]
int Engine::sample_step(int B, const StepLayout& L) {
  const hb_model_desc& d = model_.d;
  const float* temp = (const float*)(d_step_ + L.temp);
  const uint64_t* seed = (const uint64_t*)(d_step_ + L.seed);
  const int32_t* topk = (const int32_t*)(d_step_ + L.topk);
  const float* topp = (const float*)(d_step_ + L.topp);
  const bool filt = (step_flags_ & STEP_FILTER) != 0;
  if (step_flags_ & STEP_PENALTY)
    SPAN(3, 0.0, apply_penalties(stream_, logits_, d.vocab, (const int32_t*)(d_step_ + L.pen_off), d_step_ + L.pen, B, d.vocab));
  SPAN(3, 4.0 * B * d.vocab, sample_tokens(stream_, logits_, d.vocab, temp, seed, sampled_, B, d.vocab, sample_ws_,
                                           filt ? topk : nullptr, filt ? topp : nullptr));
  launches_.fetch_add(filt ? 2 : 1, std::memory_order_relaxed);  // sample_tokens is 2 kernels (+1 threshold pass)
  if (step_flags_ & STEP_LOGPROBS)
    SPAN(3, 4.0 * B * d.vocab, logprob_topk(stream_, logits_, d.vocab, d.vocab, sampled_, (const int32_t*)(d_step_ + L.lpw),
                                            lp_ids_, lp_vals_, B, HB_MAX_LOGPROBS));
  return HB_OK;
}

// Decode step (one token per sequence, B <= 256): every projection is a weight-streaming skinny GEMM whose
// fp32 partial slabs are consumed by the fused row kernel that follows it in the layer.
int Engine::forward_llama_decode(int B, const StepLayout& L) {
  if (fused_decode_ok()) return forward_llama_decode_fused(B, L);
  const hb_model_desc& d = model_.d;
  const int H = d.hidden, D = d.head_dim, QD = d.heads * D, QKV = model_.qkv_cols(), F = d.ffn;
  const int32_t* tokens = (const int32_t*)(d_step_ + L.tokens);
  const int32_t* positions = (const int32_t*)(d_step_ + L.positions);
  const int32_t* slots = (const int32_t*)(d_step_ + L.slots);
  const int32_t* ctx = (const int32_t*)(d_step_ + L.ctx);
  const int32_t* pt = (const int32_t*)(d_step_ + L.pt);
  const size_t layer_kv = (size_t)num_pages_ * d.kv_heads * page_ * D;
  auto wbytes = [&](double N, double K) { return 2.0 * N * K + 2.0 * B * K + 4.0 * B * N; };
  const double rowb = 4.0 * B * H;

  // HBM hand-over chain (kernels.h StreamSig): qkv(0) -> attn(0) -> o(0) -> gate/up(0) -> down(0) -> qkv(1) ... -> head.
  // One counter per streaming kernel of the step, zeroed first (a memset node when the step is a CUDA graph).
  static const size_t bank_bytes = [] {
    const char* e = getenv("HB_DECODE_BANK_MB");
    return (size_t)(e ? atoi(e) : 0) << 20;  // measured: no gain on the headline workload, off by default
  }();
  const bool handover = sig_ != nullptr && !profile_ && (bank_bytes > 0 || dec_trace_);
  // Dependency flags (kernels.h DepSig): inside a layer chain every kernel waits for its predecessor's release/acquire
  // counter instead of griddepcontrol.wait; the step's first kernels and the sampler keep the PDL wait.
  static const int flags_env = [] { const char* e = getenv("HB_DECODE_FLAGS"); return e ? atoi(e) : -1; }();
  const int splits = decode_splits(B);
  const bool flags = dep_ != nullptr && flags_env > 0 && splits == 1;  // measured slower than griddepcontrol.wait: opt-in
  // RoPE + KV write inside the attention kernel's prologue (one kernel and one boundary fewer per layer); needs one split
  static const int fuse_env = [] { const char* e = getenv("HB_DECODE_FUSE_ROPE"); return e ? atoi(e) : -1; }();
  const bool fuse_rope = (fuse_env > 0) && splits == 1 && !flags && page_ == 64;
  const size_t n_sig = (size_t)(5 * d.layers + 1), n_dep = (size_t)(9 * d.layers + 2);
  if (handover || flags) CU(cudaMemsetAsync(sig_, 0, (n_sig + n_dep) * sizeof(int), stream_));  // dep_ follows sig_
  int prev_idx = -1, prev_count = 0;
  int dep_next = 0, dep_prev_count = 0;
  const int* dep_prev = nullptr;
  auto next_dep = [&](int my_ctas) {  // this kernel waits for the previous kernel of the chain and owns the next counter
    DepSig dp{};
    if (!flags) return dp;
    dp.wait = dep_prev;
    dp.wait_count = dep_prev_count;
    dp.done = dep_ + dep_next++;
    dep_prev = dp.done;
    dep_prev_count = my_ctas;
    return dp;
  };
  auto next_sig = [&](int idx, int my_ctas) {
    StreamSig sg{};
    sg.dep = next_dep(my_ctas);
    if (!handover) return sg;
    if (prev_idx >= 0) { sg.wait = sig_ + prev_idx; sg.wait_count = prev_count; }
    sg.done = sig_ + idx;
    sg.bank_bytes = bank_bytes;
    sg.trace = dec_trace_ ? dec_trace_ + (size_t)idx * 16 : nullptr;
    prev_idx = idx;
    prev_count = my_ctas;
    return sg;
  };

  SPAN(3, rowb, embed_gather(stream_, tokens, model_.embed, x_, B, H));
  SPAN(3, rowb, rmsnorm(stream_, x_, model_.ll[0].attn_norm, xn_, nullptr, B, H, d.norm_eps));
  for (int l = 0; l < d.layers; ++l) {
    const LlamaLayerW& w = model_.ll[l];
    bf16* kc = kv_ + (size_t)l * 2 * layer_kv;
    bf16* vc = kc + layer_kv;
    {
      const StreamSig sg = next_sig(5 * l + 0, plan_qkv_.grid);  // layer 0: no predecessor counter -> griddepcontrol.wait
      SPAN(4, wbytes(QKV, H), gemm_skinny(stream_, plan_qkv_, xn_, H, w.wqkv, H, skinny_ws_, B, QKV, H, &sg));
    }
    if (!fuse_rope) {
      const DepSig dp = next_dep(dec_qkv_rope_ctas(B, d.heads, d.kv_heads));
      SPAN(3, 8.0 * B * QKV, dec_qkv_rope_kvwrite(stream_, skinny_ws_, plan_qkv_, qkv_, positions, slots, model_.inv_freq,
                                                   kc, vc, B, d.heads, d.kv_heads, D, page_, &dp, w.bqkv));
    }
    {
      AttnDecodeArgs a{};
      a.q = qkv_; a.ldq = QKV;
      a.k_cache = kc; a.v_cache = vc;
      a.page_table = pt; a.max_pages = max_pages_per_seq_;
      a.ctx_lens = ctx;
      a.out = attn_; a.ldo = QD;
      a.workspace = dec_ws_;
      a.B = B; a.Hq = d.heads; a.Hkv = d.kv_heads; a.D = D; a.page_size = page_;
      a.num_splits = splits;
      a.scale = 1.0f / sqrtf((float)D);
      a.num_pages = num_pages_;
      if (fuse_rope) {  // the attention CTAs build q / K / V from the QKV slabs themselves (kernels.h DecodeRope)
        a.rope.ws = skinny_ws_; a.rope.segs = plan_qkv_.seg_count; a.rope.M = B; a.rope.N = QKV; a.rope.bias = w.bqkv;
        a.rope.positions = positions; a.rope.slots = slots; a.rope.inv_freq = model_.inv_freq;
        a.rope.k_cache = kc; a.rope.v_cache = vc;
      }
      a.sig = next_sig(5 * l + 1, splits * d.kv_heads * B);
      SPAN(2, attn_bytes_, attn_decode(stream_, a));
      if (a.num_splits > 1) launches_.fetch_add(1, std::memory_order_relaxed);  // combine kernel
    }
    {
      const StreamSig sg = next_sig(5 * l + 2, plan_o_.grid);
      SPAN(4, wbytes(H, QD), gemm_skinny(stream_, plan_o_, attn_, QD, w.wo, QD, skinny_ws_, B, H, QD, &sg));
    }
    {
      const DepSig dp = next_dep(B);
      SPAN(3, 3 * rowb, dec_resid_rmsnorm(stream_, skinny_ws_, plan_o_, x_, w.mlp_norm, xn_, B, H, d.norm_eps, &dp));
    }
    {
      const StreamSig sg = next_sig(5 * l + 3, plan_gu_.grid);
      SPAN(4, wbytes(2.0 * F, H), gemm_skinny(stream_, plan_gu_, xn_, H, w.wgu, H, skinny_ws_, B, 2 * F, H, &sg));
    }
    {
      const DepSig dp = next_dep(dec_swiglu_ctas(B, F));
      SPAN(3, 10.0 * B * F, dec_swiglu(stream_, skinny_ws_, plan_gu_, h_, B, F, &dp));
    }
    {
      const StreamSig sg = next_sig(5 * l + 4, plan_down_.grid);
      SPAN(4, wbytes(H, F), gemm_skinny(stream_, plan_down_, h_, F, w.wdown, F, skinny_ws_, B, H, F, &sg));
    }
    const bf16* next_norm = (l + 1 < d.layers) ? model_.ll[l + 1].attn_norm : model_.final_norm;
    {
      const DepSig dp = next_dep(B);
      SPAN(3, 3 * rowb, dec_resid_rmsnorm(stream_, skinny_ws_, plan_down_, x_, next_norm, xn_, B, H, d.norm_eps, &dp));
    }
  }
  {
    const StreamSig sg = next_sig(5 * d.layers, plan_head_.grid);
    SPAN(4, wbytes(d.vocab, H), gemm_skinny(stream_, plan_head_, xn_, H, model_.lm_head, H, skinny_ws_, B, d.vocab, H, &sg));
  }
  SPAN(3, 8.0 * B * d.vocab, dec_sum_slabs(stream_, skinny_ws_, plan_head_, logits_, d.vocab, B, d.vocab));  // PDL wait: head GEMM complete
  return sample_step(B, L);
}

// Decode step with tile finishers (kernels.h SkinnyEpi): five kernels per layer — qkv(+RoPE/KV write), attention,
// o(+residual), gate/up(+SwiGLU), down(+residual) — no row kernel between two projections.  RMSNorm is split: the
// residual finisher writes xg = bf16(x * gain) and per-tile sums of x^2, the NEXT projection's finisher multiplies its
// rows by rstd (the GEMM is linear in the row factor).
bool Engine::fused_decode_ok() const {
  static const int env = [] { const char* e = getenv("HB_DECODE_FUSED"); return e ? atoi(e) : -1; }();  // A/B override
  const hb_model_desc& d = model_.d;
  const bool on = env >= 0 ? env != 0 : cfg_.fused_decode != 0;
  return on && cnt_qkv_ && !d.qkv_bias && (d.head_dim == 64 || d.head_dim == 128) && (d.heads * d.head_dim) % 128 == 0 &&
         (d.kv_heads * d.head_dim) % 128 == 0 && d.hidden % 128 == 0 && (2 * d.ffn) % 256 == 0;
}

int Engine::forward_llama_decode_fused(int B, const StepLayout& L) {
  const hb_model_desc& d = model_.d;
  const int H = d.hidden, D = d.head_dim, QD = d.heads * D, QKV = model_.qkv_cols(), F = d.ffn;
  const int32_t* tokens = (const int32_t*)(d_step_ + L.tokens);
  const int32_t* positions = (const int32_t*)(d_step_ + L.positions);
  const int32_t* slots = (const int32_t*)(d_step_ + L.slots);
  const int32_t* ctx = (const int32_t*)(d_step_ + L.ctx);
  const int32_t* pt = (const int32_t*)(d_step_ + L.pt);
  const size_t layer_kv = (size_t)num_pages_ * d.kv_heads * page_ * D;
  auto wbytes = [&](double N, double K) { return 2.0 * N * K + 2.0 * B * K + 4.0 * B * N; };
  const double rowb = 4.0 * B * H;
  static const size_t bank_bytes = [] {
    const char* e = getenv("HB_DECODE_BANK_MB");
    return (size_t)(e ? atoi(e) : 0) << 20;
  }();
  const bool handover = sig_ != nullptr && !profile_ && (bank_bytes > 0 || dec_trace_);
  if (handover) CU(cudaMemsetAsync(sig_, 0, (size_t)(5 * d.layers + 1) * sizeof(int), stream_));
  int prev_idx = -1, prev_count = 0;
  auto next_sig = [&](int idx, int my_ctas) {
    StreamSig sg{};
    if (!handover) return sg;
    if (prev_idx >= 0) { sg.wait = sig_ + prev_idx; sg.wait_count = prev_count; }
    sg.done = sig_ + idx;
    sg.bank_bytes = bank_bytes;
    sg.trace = dec_trace_ ? dec_trace_ + (size_t)idx * 16 : nullptr;
    prev_idx = idx;
    prev_count = my_ctas;
    return sg;
  };
  const int splits = decode_splits(B);
  const int th = H / 128;
  auto normed = [&](SkinnyEpi& e) {  // rows of the GEMM input carry gain but not rstd: the finisher applies it
    e.ss_in = ss_;
    e.ss_tiles = th;
    e.norm_h = H;
    e.eps = d.norm_eps;
  };

  SPAN(3, rowb, dec_embed_prep(stream_, tokens, model_.embed, model_.ll[0].attn_norm, x_, xn_, ss_, B, H));
  for (int l = 0; l < d.layers; ++l) {
    const LlamaLayerW& w = model_.ll[l];
    bf16* kc = kv_ + (size_t)l * 2 * layer_kv;
    bf16* vc = kc + layer_kv;
    {
      const StreamSig sg = next_sig(5 * l + 0, plan_qkv_.grid);
      SkinnyEpi e{};
      e.mode = SK_QKV_ROPE;
      e.tile_cnt = cnt_qkv_;
      normed(e);
      e.qkv_out = qkv_; e.positions = positions; e.slots = slots; e.inv_freq = model_.inv_freq;
      e.k_cache = kc; e.v_cache = vc; e.Hq = d.heads; e.Hkv = d.kv_heads; e.D = D; e.page_size = page_;
      SPAN(4, wbytes(QKV, H), gemm_skinny(stream_, plan_qkv_, xn_, H, w.wqkv, H, skinny_ws_, B, QKV, H, &sg, &e));
    }
    {
      AttnDecodeArgs a{};
      a.q = qkv_; a.ldq = QKV;
      a.k_cache = kc; a.v_cache = vc;
      a.page_table = pt; a.max_pages = max_pages_per_seq_;
      a.ctx_lens = ctx;
      a.out = attn_; a.ldo = QD;
      a.workspace = dec_ws_;
      a.B = B; a.Hq = d.heads; a.Hkv = d.kv_heads; a.D = D; a.page_size = page_;
      a.num_splits = splits;
      a.scale = 1.0f / sqrtf((float)D);
      a.num_pages = num_pages_;
      a.sig = next_sig(5 * l + 1, splits * d.kv_heads * B);
      SPAN(2, attn_bytes_, attn_decode(stream_, a));
      if (a.num_splits > 1) launches_.fetch_add(1, std::memory_order_relaxed);  // combine kernel
    }
    {
      const StreamSig sg = next_sig(5 * l + 2, plan_o_.grid);
      SkinnyEpi e{};
      e.mode = SK_RESID_NORM;
      e.tile_cnt = cnt_o_;
      e.x = x_; e.gain = w.mlp_norm; e.xg = xn_; e.ss_out = ss_;
      SPAN(4, wbytes(H, QD), gemm_skinny(stream_, plan_o_, attn_, QD, w.wo, QD, skinny_ws_, B, H, QD, &sg, &e));
    }
    {
      const StreamSig sg = next_sig(5 * l + 3, plan_gu_.grid);
      SkinnyEpi e{};
      e.mode = SK_SWIGLU;
      e.tile_cnt = cnt_gu_;
      normed(e);
      e.h = h_; e.F = F;
      SPAN(4, wbytes(2.0 * F, H), gemm_skinny(stream_, plan_gu_, xn_, H, w.wgu, H, skinny_ws_, B, 2 * F, H, &sg, &e));
    }
    {
      const StreamSig sg = next_sig(5 * l + 4, plan_down_.grid);
      SkinnyEpi e{};
      e.mode = SK_RESID_NORM;
      e.tile_cnt = cnt_down_;
      e.x = x_; e.gain = (l + 1 < d.layers) ? model_.ll[l + 1].attn_norm : model_.final_norm; e.xg = xn_; e.ss_out = ss_;
      SPAN(4, wbytes(H, F), gemm_skinny(stream_, plan_down_, h_, F, w.wdown, F, skinny_ws_, B, H, F, &sg, &e));
    }
  }
  {
    const StreamSig sg = next_sig(5 * d.layers, plan_head_.grid);
    SkinnyEpi e{};
    e.mode = SK_F32;
    e.tile_cnt = cnt_head_;
    normed(e);
    e.out_f32 = logits_; e.ldo = d.vocab;
    SPAN(4, wbytes(d.vocab, H), gemm_skinny(stream_, plan_head_, xn_, H, model_.lm_head, H, skinny_ws_, B, d.vocab, H, &sg, &e));
  }
  return sample_step(B, L);
}

int Engine::forward_bert(int T, int B, int max_seqlen, const StepLayout& L, float* d_out) {
  const hb_model_desc& d = model_.d;
  const int H = d.hidden, D = d.head_dim, F = d.ffn;
  const int32_t* tokens = (const int32_t*)(d_step_ + L.tokens);
  const int32_t* positions = (const int32_t*)(d_step_ + L.positions);
  const int32_t* cu = (const int32_t*)(d_step_ + L.cu);
  LAUNCH(bert_embed_ln(stream_, tokens, positions, model_.word, model_.pos, model_.type, model_.emb_ln_g,
                       model_.emb_ln_b, x_, T, H, d.norm_eps));
  for (int l = 0; l < d.layers; ++l) {
    const BertLayerW& w = model_.bl[l];
    {
      GemmArgs g{x_, H, w.wqkv, H, qkv_, 3 * H, nullptr, 0, w.bqkv, T, 3 * H, H, EPI_BIAS, 0};
      SPAN(0, 2.0 * T * 3 * H * H, gemm_bf16_tn(stream_, g));
    }
    {
      AttnPrefillArgs a{};
      a.q = qkv_; a.ldq = 3 * H;
      a.k = qkv_ + H; a.ldk = 3 * H;
      a.v = qkv_ + 2 * H; a.ldv = 3 * H;
      a.out = attn_; a.ldo = H;
      a.cu_seqlens = cu;
      a.B = B; a.T = T; a.max_seqlen = max_seqlen;
      a.Hq = d.heads; a.Hkv = d.heads; a.D = D;
      a.causal = 0;
      a.scale = 1.0f / sqrtf((float)D);
      SPAN(1, attn_flops_, attn_prefill(stream_, a));
    }
    {
      GemmArgs g{attn_, H, w.wo, H, xn_, H, x_, H, w.bo, T, H, H, EPI_BIAS_RESID, 0};
      SPAN(0, 2.0 * T * H * H, gemm_bf16_tn(stream_, g));
    }
    SPAN(3, 4.0 * T * H, layernorm(stream_, xn_, w.ln1_g, w.ln1_b, x_, T, H, d.norm_eps));
    {
      GemmArgs g{x_, H, w.w1, H, h_, F, nullptr, 0, w.b1, T, F, H, EPI_BIAS_GELU, 0};
      SPAN(0, 2.0 * T * F * H, gemm_bf16_tn(stream_, g));
    }
    {
      GemmArgs g{h_, F, w.w2, F, xn_, H, x_, H, w.b2, T, H, F, EPI_BIAS_RESID, 0};
      SPAN(0, 2.0 * T * H * F, gemm_bf16_tn(stream_, g));
    }
    SPAN(3, 4.0 * T * H, layernorm(stream_, xn_, w.ln2_g, w.ln2_b, x_, T, H, d.norm_eps));
  }
  LAUNCH(cls_pool_l2(stream_, x_, cu, d_out, B, H));
  return HB_OK;
}

// ------------------------------------------------------------------ prefix cache (all under mu_)
static uint64_t page_key(uint64_t parent, const int32_t* toks, int n) {
  uint64_t h = parent ^ 0x9E3779B97F4A7C15ull;
  for (int i = 0; i < n; ++i) {
    h ^= (uint64_t)(uint32_t)toks[i] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 31;
  }
  return h;
}
int32_t Engine::take_page() {
  int32_t pg;
  if (!free_pages_.empty()) {
    pg = free_pages_.back();
    free_pages_.pop_back();
  } else {
    auto it = lru_.begin();  // oldest unreferenced cached page
    pg = it->second;
    lru_.erase(it);
    cache_.erase(pmeta_[pg].key);
    pmeta_[pg].cached = false;
  }
  pmeta_[pg].ref = 1;
  return pg;
}
void Engine::drop_page(int32_t pg) {
  PageMeta& m = pmeta_[pg];
  if (--m.ref > 0) return;
  if (m.cached) {
    m.tick = ++tick_;
    lru_[m.tick] = pg;
  } else {
    free_pages_.push_back(pg);
  }
}
// Every page the sequence has completely filled (prompt or generated tokens, K/V resident) becomes addressable by content.
void Engine::register_full_pages(Request* r) {
  if (!cfg_.enable_prefix_cache) return;
  const int full = r->kv_len / page_;
  std::vector<int32_t> toks(page_);
  for (int i = r->registered; i < full; ++i) {
    for (int j = 0; j < page_; ++j) toks[j] = token_at(r, i * page_ + j);
    const uint64_t parent = r->chain_key;
    const uint64_t key = page_key(parent, toks.data(), page_);
    const int32_t pg = r->pages[i];
    PageMeta& m = pmeta_[pg];
    if (!m.cached && !cache_.count(key)) {  // same content already cached on another page: keep that one
      m.cached = true;
      m.key = key;
      m.parent = parent;
      m.toks = toks;
      cache_[key] = pg;
    }
    r->chain_key = key;
    r->registered = i + 1;
  }
}

// ------------------------------------------------------------------ scheduling
static bool wants_filter(const hb_sampling& sp, int vocab) {
  return sp.temperature > 0.f && ((sp.top_k > 0 && sp.top_k < vocab) || (sp.top_p > 0.f && sp.top_p < 1.f));
}
void Engine::finish_request(Request* r, ReqState st) {
  // mu_ held
  // last page first: the LRU then evicts a chain from its tail and the shared root (system prompt) survives longest
  for (auto it = r->pages.rbegin(); it != r->pages.rend(); ++it) drop_page(*it);
  r->pages.clear();
  r->state = st;
}

int Engine::submit(const int32_t* toks, int n, const hb_sampling* sp, uint64_t* id) {
  if (!loaded_) return fail(HB_ERR_STATE, "hb_submit before a model is loaded");
  if (model_.d.arch != HB_ARCH_LLAMA) return fail(HB_ERR_INVALID, "hb_submit needs a decoder model");
  if (!toks || n <= 0 || !sp || !id) return fail(HB_ERR_INVALID, "null/empty argument");
  if (cuda_error_.load()) return fail(HB_ERR_CUDA, "engine is in a sticky CUDA error state");
  if (n >= cfg_.max_ctx) return fail(HB_ERR_INVALID, "prompt does not fit context_length");
  for (int i = 0; i < n; ++i)
    if (toks[i] < 0 || toks[i] >= model_.d.vocab) return fail(HB_ERR_INVALID, "token id out of range");
  auto r = std::make_unique<Request>();
  r->prompt.assign(toks, toks + n);
  r->sp = *sp;
  if (r->sp.max_tokens < 1) r->sp.max_tokens = 1;
  if (n + r->sp.max_tokens > cfg_.max_ctx) r->sp.max_tokens = cfg_.max_ctx - n;
  if (r->sp.logprobs < 0 || r->sp.logprobs > HB_MAX_LOGPROBS) return fail(HB_ERR_INVALID, "logprobs must be in [0, 21]");
  if ((r->sp.presence_penalty != 0.f || r->sp.frequency_penalty != 0.f) &&
      (size_t)r->sp.max_tokens * (size_t)cfg_.max_seqs > pen_cap_)
    return fail(HB_ERR_INVALID, "max_tokens too large for a request with presence/frequency penalties on this engine");
  std::lock_guard<std::mutex> g(mu_);
  if ((int)waiting_.size() >= 65536) return fail(HB_ERR_BUSY, "queue full");
  r->id = next_id_++;
  *id = r->id;
  waiting_.push_back(r.get());
  reqs_[r->id] = std::move(r);
  cv_work_.notify_one();
  return HB_OK;
}

int Engine::poll(uint64_t id, int32_t* out, int cap, int* n_out, int* finished) {
  std::lock_guard<std::mutex> g(mu_);
  auto it = reqs_.find(id);
  if (it == reqs_.end()) return fail(HB_ERR_NOT_FOUND, "unknown request id");
  Request* r = it->second.get();
  int n = 0;
  while (r->polled < r->out.size() && n < cap) out[n++] = r->out[r->polled++];
  if (n_out) *n_out = n;
  const bool done = r->state == ReqState::FINISHED || r->state == ReqState::CANCELLED || r->state == ReqState::FAILED;
  if (finished) *finished = (done && r->polled == r->out.size()) ? (r->state == ReqState::FINISHED ? 1 : 2) : 0;
  return HB_OK;
}

int Engine::wait(uint64_t id, int timeout_ms) {
  struct Guard {  // the destructor of the engine waits for every thread parked here to leave
    std::atomic<int>& n;
    explicit Guard(std::atomic<int>& c) : n(c) { n.fetch_add(1); }
    ~Guard() { n.fetch_sub(1); }
  } guard(waiters_);
  std::unique_lock<std::mutex> g(mu_);
  if (closing_.load()) return fail(HB_ERR_STATE, "engine is shutting down");
  // the record is looked up again after every wake-up: another thread may hb_release it while this one sleeps
  auto ready = [&] {
    if (closing_.load()) return true;
    auto it = reqs_.find(id);
    if (it == reqs_.end()) return true;
    const Request* r = it->second.get();
    return r->polled < r->out.size() || r->state == ReqState::FINISHED || r->state == ReqState::CANCELLED ||
           r->state == ReqState::FAILED;
  };
  if (reqs_.find(id) == reqs_.end()) return fail(HB_ERR_NOT_FOUND, "unknown request id");
  if (timeout_ms < 0)
    cv_out_.wait(g, ready);
  else
    cv_out_.wait_for(g, std::chrono::milliseconds(timeout_ms), ready);
  if (closing_.load()) return fail(HB_ERR_STATE, "engine is shutting down");
  if (reqs_.find(id) == reqs_.end()) return fail(HB_ERR_NOT_FOUND, "request released while waiting");
  return ready() ? HB_OK : HB_ERR_BUSY;
}

int Engine::cancel(uint64_t id) {
  std::lock_guard<std::mutex> g(mu_);
  auto it = reqs_.find(id);
  if (it == reqs_.end()) return fail(HB_ERR_NOT_FOUND, "unknown request id");
  Request* r = it->second.get();
  if (r->state == ReqState::WAITING) {
    waiting_.erase(std::remove(waiting_.begin(), waiting_.end(), r), waiting_.end());
    finish_request(r, ReqState::CANCELLED);
    cv_out_.notify_all();
  } else if (r->state == ReqState::RUNNING) {
    r->cancel_flag = true;  // the step loop retires it at the next step boundary
  }
  return HB_OK;
}

int Engine::release(uint64_t id) {
  std::lock_guard<std::mutex> g(mu_);
  auto it = reqs_.find(id);
  if (it == reqs_.end()) return fail(HB_ERR_NOT_FOUND, "unknown request id");
  Request* r = it->second.get();
  if (r->state == ReqState::WAITING || r->state == ReqState::RUNNING)
    return fail(HB_ERR_STATE, "request still active; cancel it first");
  reqs_.erase(it);
  return HB_OK;
}

int Engine::captured(uint64_t id, int which, float* out, size_t cap, int* rows) {
  std::lock_guard<std::mutex> g(mu_);
  auto it = reqs_.find(id);
  if (it == reqs_.end()) return fail(HB_ERR_NOT_FOUND, "unknown request id");
  const std::vector<float>& v = (which == HB_CAPTURE_PROMPT_LOGITS) ? it->second->prompt_logits : it->second->step_logits;
  if (rows) *rows = (int)(v.size() / model_.d.vocab);
  if (out) {
    if (cap < v.size()) return fail(HB_ERR_INVALID, "capture buffer too small");
    memcpy(out, v.data(), v.size() * 4);
  }
  return HB_OK;
}

// Per-row sampler inputs of a step (mu_ not needed: the step owns its batch).  Returns the number of penalty entries.
size_t Engine::fill_sampling(Request* const* batch, int B, const StepLayout& L, int T) {
  float* temp = (float*)(h_step_ + L.temp);
  uint64_t* seed = (uint64_t*)(h_step_ + L.seed);
  int32_t* topk = (int32_t*)(h_step_ + L.topk);
  float* topp = (float*)(h_step_ + L.topp);
  int32_t* lpw = (int32_t*)(h_step_ + L.lpw);
  int32_t* pen_off = (int32_t*)(h_step_ + L.pen_off);
  struct Entry { int32_t tok; float val; };
  Entry* pen = (Entry*)(h_step_ + L.pen);
  step_flags_ = 0;
  size_t n = 0;
  for (int i = 0; i < B; ++i) {
    const Request* r = batch[i];
    temp[i] = r->sp.temperature;
    seed[i] = r->sp.seed * 0x9E3779B97F4A7C15ull + (uint64_t)r->out.size();  // one noise stream per generated position
    topk[i] = r->sp.top_k;
    topp[i] = r->sp.top_p;
    lpw[i] = std::max(0, std::min(r->sp.logprobs, (int)HB_MAX_LOGPROBS));
    if (wants_filter(r->sp, model_.d.vocab)) step_flags_ |= STEP_FILTER;
    if (lpw[i] > 0) step_flags_ |= STEP_LOGPROBS;
    pen_off[i] = (int32_t)n;
    if (r->sp.presence_penalty != 0.f || r->sp.frequency_penalty != 0.f) {
      step_flags_ |= STEP_PENALTY;
      for (const auto& kv : r->counts) {
        if (n >= pen_cap_) break;  // cannot happen: submit() bounds max_tokens of penalised requests by pen_cap_ / max_seqs
        pen[n++] = Entry{kv.first, r->sp.presence_penalty + r->sp.frequency_penalty * (float)kv.second};
      }
    }
  }
  pen_off[B] = (int32_t)n;
  (void)T;
  return n;
}

// after the step's stream sync: append each row's [width] record to its request
int Engine::collect_logprobs(Request* const* batch, int B, bool prefill) {
  if (!(step_flags_ & STEP_LOGPROBS)) return HB_OK;
  CU(cudaMemcpyAsync(h_lp_ids_, lp_ids_, (size_t)B * HB_MAX_LOGPROBS * 4, cudaMemcpyDeviceToHost, stream_));
  CU(cudaMemcpyAsync(h_lp_vals_, lp_vals_, (size_t)B * HB_MAX_LOGPROBS * 4, cudaMemcpyDeviceToHost, stream_));
  CU(cudaStreamSynchronize(stream_));
  for (int i = 0; i < B; ++i) {
    Request* r = batch[i];
    const int w = std::max(0, std::min(r->sp.logprobs, (int)HB_MAX_LOGPROBS));
    if (w == 0 || (prefill && r->prefilled + r->chunk < r->target)) continue;  // not the last chunk: nothing sampled
    r->lp_ids.insert(r->lp_ids.end(), h_lp_ids_ + (size_t)i * HB_MAX_LOGPROBS, h_lp_ids_ + (size_t)i * HB_MAX_LOGPROBS + w);
    r->lp_vals.insert(r->lp_vals.end(), h_lp_vals_ + (size_t)i * HB_MAX_LOGPROBS, h_lp_vals_ + (size_t)i * HB_MAX_LOGPROBS + w);
  }
  return HB_OK;
}

int Engine::logprobs(uint64_t id, int first_row, int max_rows, int32_t* ids, float* lps, int* rows, int* width) {
  std::lock_guard<std::mutex> g(mu_);
  auto it = reqs_.find(id);
  if (it == reqs_.end()) return fail(HB_ERR_NOT_FOUND, "unknown request id");
  const Request* r = it->second.get();
  const int w = std::max(0, std::min(r->sp.logprobs, (int)HB_MAX_LOGPROBS));
  if (width) *width = w;
  int have = w ? (int)(r->lp_ids.size() / w) : 0;
  have = std::min(have, (int)r->out.size());  // rows become visible together with their token
  if (first_row < 0) first_row = 0;
  int n = std::max(0, std::min(max_rows, have - first_row));
  if (n > 0 && ids && lps) {
    memcpy(ids, r->lp_ids.data() + (size_t)first_row * w, (size_t)n * w * 4);
    memcpy(lps, r->lp_vals.data() + (size_t)first_row * w, (size_t)n * w * 4);
  } else if (n > 0) {
    n = 0;
  }
  if (rows) *rows = n;
  return HB_OK;
}

int Engine::run_prefill(std::vector<Request*>& batch) {
  const hb_model_desc& d = model_.d;
  const int B = (int)batch.size();
  int T = 0, max_len = 0;
  bool want_all = false, paged = false;
  const int Bp = B - step_decode_rows_;  // the trailing step_decode_rows_ entries are decode rows of a mixed step
  for (int i = 0; i < B; ++i) {  // this step covers tokens [prefilled, prefilled + chunk) of every request
    Request* r = batch[i];
    T += r->chunk;
    max_len = std::max(max_len, r->chunk);
    want_all |= (r->sp.capture & HB_CAPTURE_PROMPT_LOGITS) != 0 && r->prefilled + r->chunk <= (int)r->prompt.size();
    if (i < Bp) paged |= r->prefilled > 0;
  }
  if (want_all) {
    if (T > 4096) return fail(HB_ERR_INVALID, "HB_CAPTURE_PROMPT_LOGITS limited to 4096 prompt tokens per step");
    if (all_logits_rows_ < T) {
      if (all_logits_) cudaFree(all_logits_);
      all_logits_ = nullptr;
      CU(cudaMalloc(&all_logits_, (size_t)T * d.vocab * 4));
      all_logits_rows_ = T;
    }
  }
  StepLayout L = layout(T, B);
  int32_t* tok = (int32_t*)(h_step_ + L.tokens);
  int32_t* pos = (int32_t*)(h_step_ + L.positions);
  int32_t* slot = (int32_t*)(h_step_ + L.slots);
  int32_t* cu = (int32_t*)(h_step_ + L.cu);
  int32_t* last = (int32_t*)(h_step_ + L.last);
  int32_t* ctx = (int32_t*)(h_step_ + L.ctx);
  int32_t* pt = (int32_t*)(h_step_ + L.pt);
  L.total = al16(L.pen + 8 * fill_sampling(batch.data(), B, L, T));
  int t = 0;
  for (int i = 0; i < B; ++i) {
    Request* r = batch[i];
    cu[i] = t;
    const int end = r->prefilled + r->chunk;
    for (int j = r->prefilled; j < end; ++j, ++t) {
      tok[t] = token_at(r, j);  // generated tokens too: decode rows of a mixed step, resumed (preempted) sequences
      pos[t] = j;
      slot[t] = r->pages[j / page_] * page_ + j % page_;
    }
    last[i] = t - 1;
    ctx[i] = end;
    if (paged || i >= Bp) {
      const int np = (end + page_ - 1) / page_;
      for (int j = 0; j < np; ++j) pt[(size_t)i * max_pages_per_seq_ + j] = r->pages[j];
    }
  }
  cu[B] = t;
  CU(cudaMemcpyAsync(d_step_, h_step_, L.total, cudaMemcpyHostToDevice, stream_));
  attn_flops_ = 0;
  for (Request* r : batch) {
    const double n = (double)r->chunk, s = (double)r->prefilled;
    attn_flops_ += 4.0 * (n * s + n * n * 0.5) * d.head_dim * d.heads;  // QK^T + PV over the causal trapezoid
  }
  CU(cudaEventRecord(fwd_a_, stream_));
  int rc = forward_llama(T, B, true, max_len, L, want_all, paged);
  if (rc != HB_OK) return rc;
  CU(cudaEventRecord(fwd_b_, stream_));
  CU(cudaMemcpyAsync(h_sampled_, sampled_, (size_t)B * 4, cudaMemcpyDeviceToHost, stream_));
  CU(cudaStreamSynchronize(stream_));
  {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, fwd_a_, fwd_b_) == cudaSuccess) gpu_ms_prefill_ += ms;
    drain_spans();
  }
  steps_prefill_++;
  tok_prefill_ += T;
  {
    int rc2 = collect_logprobs(batch.data(), B, true);
    if (rc2 != HB_OK) return rc2;
  }
  // captures (test tap; synchronous copies are fine here)
  for (int i = 0; i < B; ++i) {
    Request* r = batch[i];
    if ((r->sp.capture & HB_CAPTURE_PROMPT_LOGITS) && r->prefilled + r->chunk <= (int)r->prompt.size()) {
      r->prompt_logits.resize(r->prompt.size() * (size_t)d.vocab);
      CU(cudaMemcpy(r->prompt_logits.data() + (size_t)r->prefilled * d.vocab, all_logits_ + (size_t)cu[i] * d.vocab,
                    (size_t)r->chunk * d.vocab * 4, cudaMemcpyDeviceToHost));
    }
    if ((r->sp.capture & HB_CAPTURE_STEP_LOGITS) && r->prefilled + r->chunk == r->target) {
      const size_t o = r->step_logits.size();
      r->step_logits.resize(o + d.vocab);
      CU(cudaMemcpy(r->step_logits.data() + o, logits_ + (size_t)i * d.vocab, (size_t)d.vocab * 4, cudaMemcpyDeviceToHost));
    }
  }
  return HB_OK;
}

int Engine::run_decode(std::vector<Request*>& batch) {
  const hb_model_desc& d = model_.d;
  const int B = (int)batch.size();
  StepLayout L = layout(B, B);
  int32_t* tok = (int32_t*)(h_step_ + L.tokens);
  int32_t* pos = (int32_t*)(h_step_ + L.positions);
  int32_t* slot = (int32_t*)(h_step_ + L.slots);
  int32_t* cu = (int32_t*)(h_step_ + L.cu);
  int32_t* last = (int32_t*)(h_step_ + L.last);
  int32_t* ctx = (int32_t*)(h_step_ + L.ctx);
  int32_t* pt = (int32_t*)(h_step_ + L.pt);
  L.total = al16(L.pen + 8 * fill_sampling(batch.data(), B, L, B));
  for (int i = 0; i < B; ++i) {
    Request* r = batch[i];
    const int p = r->kv_len;
    tok[i] = r->out.back();
    pos[i] = p;
    slot[i] = r->pages[p / page_] * page_ + p % page_;
    cu[i] = i;
    last[i] = i;
    ctx[i] = p + 1;
    const int np = (p + 1 + page_ - 1) / page_;
    for (int j = 0; j < np; ++j) pt[(size_t)i * max_pages_per_seq_ + j] = r->pages[j];
  }
  cu[B] = B;
  CU(cudaMemcpyAsync(d_step_, h_step_, L.total, cudaMemcpyHostToDevice, stream_));
  attn_bytes_ = 0;
  for (Request* r : batch) attn_bytes_ += (double)(r->kv_len + 1) * 2.0 * d.kv_heads * d.head_dim * 2.0;
  CU(cudaEventRecord(fwd_a_, stream_));
  if (cfg_.use_cuda_graphs && !profile_) {
    const int gkey = B | (step_flags_ << 20);  // filter / penalty / log-probability passes are different kernel sequences
    auto it = graphs_.find(gkey);
    if (it == graphs_.end()) {
      cudaGraph_t graph = nullptr;
      const uint64_t l0 = launches_.load();
      CU(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal));
      int rc = forward_llama(B, B, false, 1, L, false);
      cudaError_t ce = cudaStreamEndCapture(stream_, &graph);
      graph_kernels_[gkey] = launches_.load() - l0;
      launches_.store(l0);  // captured, not executed
      if (rc != HB_OK) return rc;
      if (ce != cudaSuccess) return fail_cuda(ce, "cudaStreamEndCapture");
      cudaGraphExec_t exec = nullptr;
      CU(cudaGraphInstantiate(&exec, graph, 0));
      cudaGraphDestroy(graph);
      it = graphs_.emplace(gkey, exec).first;
    }
    CU(cudaGraphLaunch(it->second, stream_));
    graph_launches_++;
    launches_.fetch_add(graph_kernels_[gkey]);
  } else {
    int rc = forward_llama(B, B, false, 1, L, false);
    if (rc != HB_OK) return rc;
  }
  CU(cudaEventRecord(fwd_b_, stream_));
  CU(cudaMemcpyAsync(h_sampled_, sampled_, (size_t)B * 4, cudaMemcpyDeviceToHost, stream_));
  CU(cudaStreamSynchronize(stream_));
  {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, fwd_a_, fwd_b_) == cudaSuccess) gpu_ms_decode_ += ms;
    drain_spans();
  }
  steps_decode_++;
  tok_decode_ += B;
  {
    int rc2 = collect_logprobs(batch.data(), B, false);
    if (rc2 != HB_OK) return rc2;
  }
  for (int i = 0; i < B; ++i) {
    Request* r = batch[i];
    if (r->sp.capture & HB_CAPTURE_STEP_LOGITS) {
      const size_t o = r->step_logits.size();
      r->step_logits.resize(o + d.vocab);
      CU(cudaMemcpy(r->step_logits.data() + o, logits_ + (size_t)i * d.vocab, (size_t)d.vocab * 4, cudaMemcpyDeviceToHost));
    }
  }
  return HB_OK;
}

// A running sequence loses its pages (they stay content-addressed in the prefix cache while unreferenced) and goes back
// to the head of the queue; on re-admission prompt + generated tokens are prefilled again — vLLM's recompute preemption,
// what `--max-num-seqs` implies in the backend the reference spawns (api/pkg/runner/vllm_runtime.go:705-762).  mu_ held.
void Engine::preempt(Request* v) {
  running_.erase(std::remove(running_.begin(), running_.end(), v), running_.end());
  for (auto it = v->pages.rbegin(); it != v->pages.rend(); ++it) drop_page(*it);
  v->pages.clear();
  v->kv_len = v->prefilled = v->registered = 0;
  v->chain_key = 0;
  v->in_running = false;
  v->preempted += 1;
  v->state = ReqState::WAITING;
  preemptions_++;
  auto pos = waiting_.begin();
  if (pos != waiting_.end() && !(*pos)->pages.empty()) ++pos;  // never in front of a prompt that is between its chunks
  waiting_.insert(pos, v);
}

// Every running sequence owns the page its next position falls into; when the pool is empty the most recently admitted
// running sequence is preempted (possibly the one that asked).  mu_ held.
void Engine::ensure_decode_pages() {
  for (size_t i = 0; i < running_.size();) {
    Request* r = running_[i];
    const int need_idx = r->kv_len / page_;
    bool gone = false;
    while ((int)r->pages.size() <= need_idx) {
      if (pages_available() > 0) {
        r->pages.push_back(take_page());
      } else {
        Request* v = running_.back();
        preempt(v);
        if (v == r) { gone = true; break; }
      }
    }
    if (!gone) ++i;
  }
}

int Engine::step(int* did_work) {
  if (did_work) *did_work = 0;
  if (!loaded_) return fail(HB_ERR_STATE, "hb_step before a model is loaded");
  if (model_.d.arch != HB_ARCH_LLAMA) return HB_OK;
  if (cuda_error_.load()) return fail(HB_ERR_CUDA, "engine is in a sticky CUDA error state");
  std::lock_guard<std::mutex> gg(gpu_mu_);
  CU(cudaSetDevice(cfg_.device));
  SmLimitScope sm_scope(cfg_.sm_budget);
  std::vector<Request*> batch;
  bool prefill = false;
  {
    std::lock_guard<std::mutex> g(mu_);
    // retire cancelled sequences
    for (size_t i = 0; i < running_.size();) {
      if (running_[i]->cancel_flag) {
        running_[i]->in_running = false;
        finish_request(running_[i], ReqState::CANCELLED);
        running_.erase(running_.begin() + i);
        cv_out_.notify_all();
      } else {
        ++i;
      }
    }
    if (!waiting_.empty() && waiting_.front()->cancel_flag) {  // a partly prefilled prompt cancelled between its chunks
      finish_request(waiting_.front(), ReqState::CANCELLED);
      waiting_.pop_front();
      cv_out_.notify_all();
    }
    // Token budget of this step.  With decode_with_prefill the running sequences ride along in prefill steps (one token
    // each, attending to their cached context through the paged attention path — vLLM's mixed batches), and such steps
    // are kept short (mixed_step_tokens) so that a long prompt is spread over several of them instead of stalling every
    // running stream for the length of a full prefill step.
    const bool mixed = cfg_.decode_with_prefill != 0 && !running_.empty();
    const int mixed_cap = cfg_.mixed_step_tokens > 0 ? cfg_.mixed_step_tokens : 2048;
    const int budget = mixed ? std::max(64, std::min(t_cap_, mixed_cap) - (int)running_.size()) : t_cap_;
    // admission: FIFO.  A sequence takes pages for what it is about to prefill plus its first generated token; further
    // pages are taken as it grows (ensure_decode_pages), so max_tokens does not hold memory it may never use.  Prompts are
    // packed whole into the step's token budget; only a prompt LONGER than the budget is split, and then runs as
    // budget-sized chunks (chunks after the first attend to the already cached prefix through the paged pool).
    int T = 0;
    std::vector<int32_t> seq;
    while (!waiting_.empty()) {
      Request* r = waiting_.front();
      if (r->pages.empty()) {
        // a preempted sequence resumes by prefilling prompt + everything it had generated
        const int n = (int)(r->prompt.size() + r->out.size());
        const int32_t* toks = r->prompt.data();
        if (!r->out.empty()) {
          seq.assign(r->prompt.begin(), r->prompt.end());
          seq.insert(seq.end(), r->out.begin(), r->out.end());
          toks = seq.data();
        }
        // positions never reach max_ctx (generation stops there), so a sequence never needs more than a full context of pages
        const int total = std::min((n + 1 + page_ - 1) / page_, max_pages_per_seq_);
        if ((int)(running_.size() + batch.size()) >= cfg_.max_seqs) break;
        // leading full pages already in the pool (always leave >= 1 token to run: its logits seed the decode)
        std::vector<int32_t> hit;
        uint64_t key = 0;
        int idle_hits = 0;
        if (cfg_.enable_prefix_cache) {
          for (int i = 0; (i + 1) * page_ <= n - 1; ++i) {
            const uint64_t k2 = page_key(key, toks + (size_t)i * page_, page_);
            auto it = cache_.find(k2);
            if (it == cache_.end()) break;
            const PageMeta& m = pmeta_[it->second];
            if (m.parent != key || !std::equal(m.toks.begin(), m.toks.end(), toks + (size_t)i * page_)) break;
            hit.push_back(it->second);
            idle_hits += m.ref == 0;
            key = k2;
          }
        }
        const int need = total - (int)hit.size();
        const int rest = n - (int)hit.size() * page_;  // tokens still to prefill
        // one spare page per running sequence stays free: admitting into the last pages would only force a preemption
        if (pages_available() - idle_hits < need + (int)running_.size()) break;
        if (rest <= budget && T + rest > budget) break;
        if (rest > budget && T > 0) break;
        for (int32_t pg : hit) {
          PageMeta& m = pmeta_[pg];
          if (m.ref++ == 0) lru_.erase(m.tick);
          r->pages.push_back(pg);
        }
        for (int i = 0; i < need; ++i) r->pages.push_back(take_page());
        r->prefilled = r->kv_len = (int)hit.size() * page_;
        r->registered = (int)hit.size();
        r->chain_key = key;
        r->target = n;
        prefix_hit_tokens_ += (uint64_t)r->prefilled;
        r->state = ReqState::RUNNING;  // owns cache pages from here on: cancellation goes through cancel_flag
      }
      r->chunk = std::min(r->target - r->prefilled, budget - T);
      if (r->chunk <= 0) break;
      batch.push_back(r);
      T += r->chunk;
      if (r->prefilled + r->chunk < r->target) break;  // stays at the head of the queue until its last chunk
      waiting_.pop_front();
    }
    ensure_decode_pages();
    step_decode_rows_ = 0;
    if (!batch.empty()) {
      prefill = true;
      if (cfg_.decode_with_prefill) {
        step_decode_rows_ = (int)running_.size();
        for (Request* r : running_) {  // decode rows: a one-token chunk at the end of the cached sequence
          r->prefilled = r->kv_len;
          r->chunk = 1;
          r->target = r->kv_len + 1;
          batch.push_back(r);
        }
        if (!running_.empty()) steps_mixed_++;
      }
    } else {
      batch = running_;
    }
  }
  if (batch.empty()) return HB_OK;
  int rc = prefill ? run_prefill(batch) : run_decode(batch);
  {
    std::lock_guard<std::mutex> g(mu_);
    if (rc != HB_OK) {
      for (Request* r : batch) {
        running_.erase(std::remove(running_.begin(), running_.end(), r), running_.end());
        waiting_.erase(std::remove(waiting_.begin(), waiting_.end(), r), waiting_.end());  // a partly prefilled prompt
        r->in_running = false;
        finish_request(r, ReqState::FAILED);
      }
      cv_out_.notify_all();
      return rc;
    }
    for (int i = 0; i < (int)batch.size(); ++i) {
      Request* r = batch[i];
      const int32_t t = h_sampled_[i];
      if (prefill) {
        r->prefilled += r->chunk;
        r->kv_len = r->prefilled;
        register_full_pages(r);
        if (r->prefilled < r->target) continue;  // more chunks to go: nothing sampled yet
        if (!r->in_running) {
          running_.push_back(r);
          r->in_running = true;
        }
      } else {
        r->kv_len += 1;
        register_full_pages(r);
      }
      r->out.push_back(t);
      if (r->sp.presence_penalty != 0.f || r->sp.frequency_penalty != 0.f) r->counts[t] += 1;
      const bool done = (int)r->out.size() >= r->sp.max_tokens || (r->sp.eos_token >= 0 && t == r->sp.eos_token) ||
                        r->kv_len + 1 >= cfg_.max_ctx;
      if (done) {
        running_.erase(std::remove(running_.begin(), running_.end(), r), running_.end());
        r->in_running = false;
        finish_request(r, ReqState::FINISHED);
      }
    }
    cv_out_.notify_all();
  }
  if (did_work) *did_work = 1;
  return HB_OK;
}

void Engine::loop() {
  while (!stop_.load()) {
    {
      std::unique_lock<std::mutex> g(mu_);
      cv_work_.wait_for(g, std::chrono::milliseconds(50),
                        [&] { return stop_.load() || !waiting_.empty() || !running_.empty(); });
      if (stop_.load()) break;
      if (waiting_.empty() && running_.empty()) continue;
      if (running_.empty() && !waiting_.empty() && waiting_.front()->pages.empty()) {
        // Idle engine woken by the first request of what is usually a burst (the scheduler releases queued work in one
        // go, api/pkg/scheduler/scheduler.go:1378-1420): give the rest a moment to arrive (until 150 us pass without a new one, 2 ms at most) so the first prefill step
        // is a full one instead of a single prompt.  Costs an idle engine at most that much time-to-first-token.
        size_t seen = waiting_.size();
        const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(2);
        while ((int)seen < cfg_.max_seqs && !stop_.load() && std::chrono::steady_clock::now() < deadline) {
          cv_work_.wait_for(g, std::chrono::microseconds(150));  // every submission notifies: returns early while they keep coming
          if (waiting_.size() == seen) break;                     // a quiet 150 us: the burst is over
          seen = waiting_.size();
        }
        if (stop_.load()) break;
      }
    }
    int did = 0;
    int rc = step(&did);
    if (rc != HB_OK && rc != HB_ERR_STATE) {
      // sticky CUDA failure: fail everything still queued so callers do not hang
      std::lock_guard<std::mutex> g(mu_);
      for (Request* r : waiting_) finish_request(r, ReqState::FAILED);
      waiting_.clear();
      for (Request* r : running_) finish_request(r, ReqState::FAILED);
      running_.clear();
      cv_out_.notify_all();
      if (rc == HB_ERR_CUDA) break;
    }
    if (!did) std::this_thread::sleep_for(std::chrono::microseconds(200));  // admission blocked (no pages / slots)
  }
}

int Engine::start() {
  if (!loaded_) return fail(HB_ERR_STATE, "hb_engine_start before a model is loaded");
  if (thread_running_) return HB_OK;
  stop_.store(false);
  thread_ = std::thread([this] { loop(); });
  thread_running_ = true;
  return HB_OK;
}

int Engine::stop() {
  if (!thread_running_) return HB_OK;
  stop_.store(true);
  cv_work_.notify_all();
  if (thread_.joinable()) thread_.join();
  thread_running_ = false;
  return HB_OK;
}

// ------------------------------------------------------------------ embeddings (BERT-style encoders)
int Engine::embed(const int32_t* toks, const int32_t* offsets, int nseq, float* out) {
  if (!loaded_) return fail(HB_ERR_STATE, "hb_embed before a model is loaded");
  if (nseq < 0 || (nseq > 0 && (!toks || !offsets || !out))) return fail(HB_ERR_INVALID, "null argument");
  if (cuda_error_.load()) return fail(HB_ERR_CUDA, "engine is in a sticky CUDA error state");
  const hb_model_desc& d = model_.d;
  const bool dec = d.arch == HB_ARCH_LLAMA;  // decoder embedder: causal pass without KV-cache writes, last-token pooling
  for (int i = 0; i < nseq; ++i) {
    const int n = offsets[i + 1] - offsets[i];
    if (n <= 0) return fail(HB_ERR_INVALID, "empty sequence");
    if (n > cfg_.max_ctx || n > d.max_pos || n > t_cap_) return fail(HB_ERR_INVALID, "sequence longer than the model's positions");
  }
  for (int i = (nseq ? offsets[0] : 0); i < (nseq ? offsets[nseq] : 0); ++i)
    if (toks[i] < 0 || toks[i] >= d.vocab) return fail(HB_ERR_INVALID, "token id out of range");
  std::lock_guard<std::mutex> gg(gpu_mu_);
  CU(cudaSetDevice(cfg_.device));
  SmLimitScope sm_scope(cfg_.sm_budget);
  step_decode_rows_ = 0;  // an embedding pass has no decode rows (the field belongs to the last scheduler step)
  const int bmax = dec ? cfg_.max_seqs : b_cap_;
  int s0 = 0;
  while (s0 < nseq) {
    int s1 = s0, T = 0, max_len = 0;
    while (s1 < nseq && s1 - s0 < bmax) {
      const int n = offsets[s1 + 1] - offsets[s1];
      if (T + n > t_cap_) break;
      T += n;
      max_len = std::max(max_len, n);
      ++s1;
    }
    const int B = s1 - s0;
    const StepLayout L = layout(T, B);
    int32_t* tok = (int32_t*)(h_step_ + L.tokens);
    int32_t* pos = (int32_t*)(h_step_ + L.positions);
    int32_t* cu = (int32_t*)(h_step_ + L.cu);
    const int base = offsets[s0];
    memcpy(tok, toks + base, (size_t)T * 4);
    for (int i = 0; i < B; ++i) {
      const int o = offsets[s0 + i] - base, n = offsets[s0 + i + 1] - offsets[s0 + i];
      cu[i] = o;
      for (int j = 0; j < n; ++j) pos[o + j] = j;
    }
    cu[B] = T;
    if (dec) {
      int32_t* slot = (int32_t*)(h_step_ + L.slots);
      int32_t* last = (int32_t*)(h_step_ + L.last);
      int32_t* ctx = (int32_t*)(h_step_ + L.ctx);
      for (int t = 0; t < T; ++t) slot[t] = -1;  // nothing enters the paged pool
      for (int i = 0; i < B; ++i) {
        last[i] = cu[i + 1] - 1;
        ctx[i] = i;
      }
    }
    CU(cudaMemcpyAsync(d_step_, h_step_, L.total, cudaMemcpyHostToDevice, stream_));
    attn_flops_ = 0;
    for (int i = 0; i < B; ++i) {
      const double n = offsets[s0 + i + 1] - offsets[s0 + i];
      attn_flops_ += (dec ? 2.0 : 4.0) * n * n * d.head_dim * d.heads;
    }
    CU(cudaEventRecord(fwd_a_, stream_));
    int rc = dec ? forward_llama(T, B, true, max_len, L, false, false, d_embed_out_) : forward_bert(T, B, max_len, L, d_embed_out_);
    if (rc != HB_OK) return rc;
    CU(cudaEventRecord(fwd_b_, stream_));
    CU(cudaMemcpyAsync(h_embed_out_, d_embed_out_, (size_t)B * d.hidden * 4, cudaMemcpyDeviceToHost, stream_));
    CU(cudaStreamSynchronize(stream_));
    {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, fwd_a_, fwd_b_) == cudaSuccess) gpu_ms_prefill_ += ms;
      drain_spans();
    }
    memcpy(out + (size_t)s0 * d.hidden, h_embed_out_, (size_t)B * d.hidden * 4);
    steps_prefill_++;
    tok_prefill_ += T;
    s0 = s1;
  }
  return HB_OK;
}

int Engine::stats(hb_stats* s) {
  if (!s) return fail(HB_ERR_INVALID, "null stats");
  if (dec_trace_ && getenv("HB_DEC_TRACE_DUMP")) {  // debug: timeline of the LAST decode step, ns relative to its first event
    std::lock_guard<std::mutex> gg(gpu_mu_);
    const int n = 5 * model_.d.layers + 1;
    std::vector<unsigned long long> h((size_t)n * 16);
    cudaMemcpy(h.data(), dec_trace_, h.size() * 8, cudaMemcpyDeviceToHost);
    unsigned long long t0 = ~0ull;
    for (auto v : h) if (v && v < t0) t0 = v;
    FILE* f = fopen(getenv("HB_DEC_TRACE_DUMP"), "w");
    if (f) {
      static const char* names[5] = {"qkv", "attn", "o", "gu", "down"};
      for (int i = 0; i < n; ++i) {
        fprintf(f, "%3d %-5s", i / 5, i == n - 1 ? "head" : names[i % 5]);
        for (int e = 0; e < 8; ++e) fprintf(f, " %9lld", h[(size_t)i * 16 + e] ? (long long)(h[(size_t)i * 16 + e] - t0) : -1ll);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }
  memset(s, 0, sizeof *s);
  std::lock_guard<std::mutex> g(mu_);
  s->weights_bytes = model_.arena_bytes;
  s->kv_bytes = kv_bytes_;
  s->workspace_bytes = ws_bytes_ + step_bytes_;
  s->budget_bytes = budget_;
  s->kv_pages_total = num_pages_;
  s->kv_pages_free = pages_available();
  s->kv_pages_cached = (int)lru_.size();
  s->prefix_hit_tokens = prefix_hit_tokens_;
  s->preemptions = (int32_t)preemptions_;
  s->steps_mixed = steps_mixed_;
  s->running = (int)running_.size();
  s->waiting = (int)waiting_.size();
  s->steps_prefill = steps_prefill_;
  s->steps_decode = steps_decode_;
  s->tokens_prefill = tok_prefill_;
  s->tokens_decode = tok_decode_;
  s->kernel_launches = launches_;
  s->graph_launches = graph_launches_;
  s->cuda_error = cuda_error_.load();
  s->gpu_ms_prefill = gpu_ms_prefill_;
  s->gpu_ms_decode = gpu_ms_decode_;
  for (int i = 0; i < 8; ++i) {
    s->prof_ms[i] = prof_ms_[i];
    s->prof_work[i] = prof_work_[i];
    s->prof_launches[i] = prof_launches_[i];
  }
  return HB_OK;
}

}  // namespace hb
