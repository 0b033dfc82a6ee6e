/* helix_b200.h — C ABI of the B200-native runtime that replaces the child-process backends
 * (ollama serve / vllm api_server) behind Helix's runner.Runtime interface.
 *
 * What each entry point replaces in the reference (helixml/helix @ 2a205c7):
 *   hb_engine_create / hb_model_load_* / hb_engine_start
 *        <- Runtime.Start: OllamaRuntime.Start (api/pkg/runner/ollama_runtime.go:179-276) and
 *           VLLMRuntime.Start (api/pkg/runner/vllm_runtime.go:163-252): spawn backend on gpu_index,
 *           load weights, size the KV pool from --gpu-memory-utilization / ModelMemoryRequirement,
 *           --max-num-seqs, --max-model-len (api/pkg/scheduler/runner.go:1187-1259,1344-1397).
 *   hb_engine_destroy
 *        <- Runtime.Stop (api/pkg/runner/slot.go:113-140): must release ALL device memory synchronously.
 *   hb_submit / hb_poll / hb_cancel
 *        <- the HTTP hop createChatCompletion makes to slot.URL()+"/v1/chat/completions"
 *           (api/pkg/runner/openai_chat_handlers.go:100-175): token ids in, sampled token ids out,
 *           one poll per SSE chunk; finish reason closes the stream.
 *   hb_embed
 *        <- createEmbedding's proxy to "/v1/embeddings" (api/pkg/runner/openai_embedding_handlers.go:569).
 *   hb_stats / hb_last_error
 *        <- Runtime.Status (slot.go:46-57): non-empty status == running (scheduler/scheduler.go:940).
 *   hb_memory_estimate
 *        <- POST /api/v1/memory-estimate (api/pkg/runner/memory_estimation_handlers.go:36-327).
 *   hb_slot_config
 *        <- Slot.Create's decoding of CreateRunnerSlotAttributes / runtime_args (api/pkg/runner/slot.go:395-470).
 *   hb_logprobs, hb_sampling.{logprobs, presence_penalty, frequency_penalty}
 *        <- request fields the runner forwards untouched (openai.ChatCompletionRequest, openai_chat_handlers.go:100-175).
 *   hb_replica_unique_id / hb_model_load_broadcast
 *        <- the backend's own NCCL when a model spans the box's GPUs (api/pkg/runner/vllm_runtime.go:748-754); here:
 *           replicas, weights copied by one broadcast.
 *   hb_model_load_gguf / hb_gguf_describe
 *        <- llama.cpp's GGUF load inside `ollama serve` (api/pkg/runner/ollama_runtime.go:546-697).
 *   hb_tok_*
 *        <- the tokenizer + chat template the backends apply inside their child process (Dockerfile.runner:61-74,188-191).
 *
 * Conventions: 0 = ok, negative = hb_status error; caller owns every host buffer; the engine owns all
 * device memory and (after hb_engine_start) one step-loop thread; all entry points are thread-safe; no
 * callbacks into the caller (cgo-friendly). There is NO CPU fallback: without a CUDA device every
 * entry point that needs one fails with HB_ERR_CUDA.
 */
#ifndef HELIX_B200_H_
#define HELIX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_ABI_VERSION 2

typedef enum hb_status {
  HB_OK = 0,
  HB_ERR_INVALID = -1,   /* bad argument / unsupported shape */
  HB_ERR_CUDA = -2,      /* CUDA runtime error (sticky: engine must be destroyed) */
  HB_ERR_OOM = -3,       /* does not fit the memory budget */
  HB_ERR_STATE = -4,     /* call order violated (e.g. submit before model load) */
  HB_ERR_NOT_FOUND = -5, /* unknown request id / tensor name */
  HB_ERR_BUSY = -6,      /* queue full / no KV pages: retry later */
} hb_status;

typedef enum hb_arch { HB_ARCH_LLAMA = 0, HB_ARCH_BERT = 1 } hb_arch;

typedef struct hb_engine hb_engine;

typedef struct hb_engine_cfg {
  int32_t device;               /* CUDA ordinal == CreateRunnerSlotAttributes.gpu_index (types/runner.go:92-104) */
  uint64_t memory_budget_bytes; /* == model_memory_requirement: weights + KV pool + workspace must fit; 0 = all free */
  int32_t max_seqs;             /* --max-num-seqs (default 256, types/memory.go:11) */
  int32_t max_ctx;              /* context_length / --max-model-len */
  int32_t max_batched_tokens;   /* prompt tokens processed per prefill step (default 16384) */
  int32_t kv_page_size;         /* tokens per KV page; must be 64 */
  int32_t use_cuda_graphs;      /* capture the decode step per batch size */
  int32_t enable_prefix_cache;  /* --enable-prefix-caching (a vLLM arg the slot's runtime_args may carry,
                                   api/pkg/runner/vllm_runtime.go:705-762): full KV pages (64 tokens) are content-addressed; a new prompt
                                   whose leading pages are already in the pool (an earlier turn of the same chat, a
                                   shared system prompt) only prefills the rest */
  int32_t sm_budget;            /* SMs this ModelInstance may occupy (0 = the whole device).  Multi-model packing: the
                                   scheduler places several slots on one GPU by memory (api/pkg/scheduler/global_allocator.go:349-452);
                                   every persistent kernel of this engine sizes its grid to sm_budget so co-resident
                                   engines run side by side on disjoint SMs instead of time-slicing whole kernels.  With
                                   sm_partition != 0 the budget is also enforced by the hardware (CUDA green context). */
  int32_t sm_partition;         /* 1: create the engine's stream inside a green context of sm_budget SMs */
  int32_t stream_priority;      /* 0 default; 1 = high priority stream (small latency-bound models in a pack) */
  int32_t decode_with_prefill;  /* 1 (what the runtime passes): running sequences decode inside prefill steps (vLLM-style
                                   mixed batches) so a long prompt never stalls the streams; 0: prefill steps exclude
                                   decode rows (benchmark-pure phases) */
  int32_t fused_decode;         /* 1: decode GEMMs run with tile finishers (the last CTA of each stream-K tile applies RoPE+KV
                                   write / residual+norm statistics / SwiGLU itself: 5 kernels per layer instead of 9).
                                   Measured slower than the separate row kernels on B200, so off by default */
  int32_t mixed_step_tokens;    /* token budget of a step that carries decode rows (decode_with_prefill): a long prompt
                                   is spread over steps of this size so inter-token latency stays bounded; 0 = 2048 */
} hb_engine_cfg;

typedef struct hb_model_desc {
  int32_t arch;        /* hb_arch */
  int32_t hidden, layers, heads, kv_heads, head_dim, ffn, vocab;
  int32_t max_pos;     /* BERT position rows / Llama max positions */
  int32_t type_vocab;  /* BERT token-type rows (row 0 is used) */
  int32_t tie_embeddings;
  float norm_eps;
  float rope_theta;
  float rope_factor;   /* llama3 rope scaling; <= 0 disables */
  float rope_low_freq_factor, rope_high_freq_factor;
  int32_t rope_orig_max_pos;
  int32_t qkv_bias;    /* 1: q/k/v projections carry biases (Qwen2-family decoders, e.g. the reference's default embedding
                          model MrLight/dse-qwen2-2b-mrl-v1, api/pkg/model/models.go:421-433); everything else as Llama */
  int32_t reserved[7];
} hb_model_desc;

typedef struct hb_sampling {
  float temperature;   /* <= 0: greedy argmax. (The runner forces 0.1 when a request carries 0:
                          api/pkg/runner/openai_chat_handlers.go:52-58 — that policy stays in the host shim.) */
  uint64_t seed;
  int32_t max_tokens;  /* generated tokens, >= 1 */
  int32_t eos_token;   /* < 0: none */
  int32_t capture;     /* HB_CAPTURE_* bit mask (parity tap) */
  int32_t top_k;       /* sampled rows only: keep the k most likely tokens (ties at the k-th logit kept); <= 0: off.
                          top_k / top_p are request fields the reference forwards untouched to its backend
                          (openai.ChatCompletionRequest, api/pkg/runner/openai_chat_handlers.go:100-175) */
  float top_p;         /* nucleus: smallest set of most likely tokens (after top_k) with softmax(logits/T) mass >= top_p;
                          values outside (0,1) (so also a zero-initialised struct): off */
  int32_t logprobs;    /* 0: off.  n >= 1: record, for every generated token, the log-probability of the chosen token and of
                          the n-1 most likely alternatives (OpenAI `logprobs: true, top_logprobs: n-1`; n <= 21), read back
                          with hb_logprobs.  Computed on the GPU from the penalised, un-tempered logits. */
  float presence_penalty;  /* OpenAI presence_penalty: subtracted once from the logit of every token already generated */
  float frequency_penalty; /* OpenAI frequency_penalty: subtracted per previous occurrence of the token in the output */
  int32_t reserved2[4];
} hb_sampling;
#define HB_MAX_LOGPROBS 21

#define HB_CAPTURE_NONE 0
#define HB_CAPTURE_STEP_LOGITS 1   /* fp32 logits row of every generated token */
#define HB_CAPTURE_PROMPT_LOGITS 2 /* fp32 logits of every prompt position */

typedef struct hb_stats {
  uint64_t weights_bytes, kv_bytes, workspace_bytes, budget_bytes;
  int32_t kv_pages_total, kv_pages_free;
  int32_t running, waiting;
  uint64_t steps_prefill, steps_decode;
  uint64_t tokens_prefill, tokens_decode;
  uint64_t kernel_launches;   /* launches of this library's kernels since creation */
  uint64_t graph_launches;
  int32_t cuda_error;         /* sticky cudaError_t, 0 = healthy */
  int32_t kv_pages_cached;    /* unreferenced pages kept for prefix reuse (counted in kv_pages_free: evictable) */
  int32_t preemptions;        /* running sequences evicted for pages and re-queued (KV recomputed on re-admission) */
  uint64_t prefix_hit_tokens; /* prompt tokens served from cached pages instead of being prefilled */
  uint64_t steps_mixed;       /* prefill steps that also carried decode rows of running sequences */
  int32_t reserved[1];
  /* device time of the forward passes (CUDA events on the engine stream: after the step's inputs are
     resident, before the sampled ids are copied back) */
  double gpu_ms_prefill, gpu_ms_decode;
  /* hb_set_profile(e,1): per-launch CUDA-event spans by kernel family
     [0]=tcgen05 GEMM in prefill steps, [1]=prefill attention, [2]=decode attention, [3]=row kernels,
     [4]=tcgen05 GEMM in decode steps */
  double prof_ms[8];
  double prof_work[8];        /* algorithmic FLOPs ([0],[1]) or bytes ([2],[3],[4]) of the timed launches */
  uint64_t prof_launches[8];
} hb_stats;

/* ---- slot creation: the scheduler's JSON -> engine configuration.
 * `json` is a types.CreateRunnerSlotRequest or its `attributes` object (api/pkg/types/runner.go:92-109) with
 * runtime "vllm" (plug-in option A).  Follows Slot.Create (api/pkg/runner/slot.go:395-470: runtime_args.model
 * overrides model; runtime_args.args as a string array, a mixed array or a {flag: value} map) and reads the flags the
 * scheduler emits (api/pkg/scheduler/runner.go:1187-1259,1344-1397): --gpu-memory-utilization, --max-num-seqs (default 256),
 * --max-model-len, --task embed, --max-num-batched-tokens, --[no-]enable-prefix-caching.  memory_budget_bytes is
 * model_memory_requirement when present, else utilization x per_gpu_memory_bytes.  max_ctx 0 = the model's default. ---- */
typedef struct hb_slot_info {
  char model[256];
  int32_t is_embed;             /* --task embed */
  int32_t tensor_parallel_size; /* as sent; this runtime replicates instead of splitting */
  int32_t n_unknown_args;       /* flags the engine has no use for (ignored, counted) */
  float gpu_memory_utilization; /* as sent (0 = absent) */
} hb_slot_info;
int hb_slot_config(const char* json, uint64_t per_gpu_memory_bytes, hb_engine_cfg* cfg_out, hb_slot_info* info_out);

/* ---- lifecycle ---- */
int hb_abi_version(void);
int hb_engine_create(const hb_engine_cfg* cfg, hb_engine** out);
void hb_engine_destroy(hb_engine* e);
const char* hb_last_error(hb_engine* e); /* e may be NULL: last create error of this thread */

/* ---- model load: describe, upload tensors by their HF checkpoint names (bf16, host memory), finish.
 *      hb_model_load_random fills the same arena on the device from `seed` (benchmarks, NCCL-root). ---- */
int hb_model_load_begin(hb_engine* e, const hb_model_desc* desc);
int hb_model_tensor_set(hb_engine* e, const char* name, const void* host_bf16, size_t n_elems);
int hb_model_load_finish(hb_engine* e);
int hb_model_load_random(hb_engine* e, const hb_model_desc* desc, uint64_t seed);
/* ---- GGUF (llama.cpp / Ollama blobs, the format of the reference's default catalogue, e.g. "llama3:instruct" = 8B Q4_0:
 *      api/pkg/model/models.go:259-266): metadata -> description, tensors dequantised to bf16 (F32 F16 BF16 Q8_0 Q4_0 Q4_1
 *      Q5_0 Q5_1 Q4_K Q5_K Q6_K), llama.cpp's q/k row permutation undone.  hb_gguf_read_tensor returns one tensor as fp32 in
 *      the HF layout (tests, tooling); HB_ERR_BUSY with *rows / *cols set when `cap` floats are not enough. ---- */
int hb_gguf_describe(const char* path, hb_model_desc* desc_out);
int hb_model_load_gguf(hb_engine* e, const char* path);
int hb_gguf_read_tensor(const char* path, const char* hf_name, float* out, size_t cap_floats, size_t* rows, size_t* cols);
/* device address/size of the contiguous weight arena: replicas receive it by one NCCL broadcast */
int hb_model_weights_arena(hb_engine* e, void** dev_ptr, size_t* bytes);
/* ---- replicas: one engine per GPU, weights loaded on rank 0 and copied to the others by ONE
 *      ncclBroadcast of the arena over NVLink; nothing collective on the request path.  The reference's only multi-GPU
 *      mechanism is the backend's own NCCL (api/pkg/runner/vllm_runtime.go:748-754); the Go host needs no NCCL binding:
 *      it moves the 128-byte id between its per-GPU runtimes (same process) or runners.
 *      hb_replica_unique_id: ncclGetUniqueId into id[HB_REPLICA_ID_BYTES] (call once, on the root).
 *      hb_model_load_broadcast: every rank calls it with the same id/world.  rank 0 must already hold a loaded model
 *      (hb_model_load_finish / hb_model_load_random) and sends; the other ranks must be freshly created engines: they
 *      allocate the arena for `desc`, receive it and finish loading.  *seconds (optional) = duration of the broadcast
 *      itself (communicator setup excluded).  libnccl.so.2 is opened on first use; HB_ERR_STATE if it is not installed. */
#define HB_REPLICA_ID_BYTES 128
int hb_replica_unique_id(void* id);
int hb_model_load_broadcast(hb_engine* e, const hb_model_desc* desc, const void* id, int32_t rank, int32_t world,
                            double* seconds);
/* closed-form footprint for the scheduler's packing (weights + max_seqs*max_ctx KV + workspace) */
int hb_memory_estimate(const hb_model_desc* desc, const hb_engine_cfg* cfg, uint64_t* weights, uint64_t* kv,
                       uint64_t* workspace);

/* ---- generation ---- */
int hb_engine_start(hb_engine* e); /* spawn the step-loop thread (continuous batching) */
int hb_engine_stop(hb_engine* e);
/* change hb_engine_cfg.decode_with_prefill / mixed_step_tokens of a live engine (takes effect at the next step) */
int hb_engine_set_mixed(hb_engine* e, int32_t decode_with_prefill, int32_t mixed_step_tokens);
int hb_step(hb_engine* e, int* did_work); /* run ONE scheduler step on the caller's thread (no step-loop thread) */
int hb_submit(hb_engine* e, const int32_t* tokens, int32_t n_tokens, const hb_sampling* sp, uint64_t* req_id);
int hb_poll(hb_engine* e, uint64_t req_id, int32_t* out_tokens, int32_t cap, int32_t* n_out, int32_t* finished);
int hb_wait(hb_engine* e, uint64_t req_id, int32_t timeout_ms); /* block until new tokens / finished */
int hb_cancel(hb_engine* e, uint64_t req_id);
int hb_release(hb_engine* e, uint64_t req_id); /* drop a finished request's record */
/* parity tap: rows captured for req (see HB_CAPTURE_*), fp32 [rows][vocab] */
int hb_captured_logits(hb_engine* e, uint64_t req_id, int32_t which, float* out, size_t cap_floats, int32_t* rows);

/* rows [first_row, first_row + max_rows) of the request's log-probability record (hb_sampling.logprobs = width):
 * ids / logprobs are [rows][width], column 0 = the sampled token, columns 1.. = the most likely tokens in descending
 * order.  *rows = rows written, *width = the request's width.  <- `logprobs` / `top_logprobs` of
 * openai.ChatCompletionRequest, forwarded untouched by api/pkg/runner/openai_chat_handlers.go:100-175. */
int hb_logprobs(hb_engine* e, uint64_t req_id, int32_t first_row, int32_t max_rows, int32_t* ids, float* logprobs,
                int32_t* rows, int32_t* width);

/* ---- embeddings: nseq sequences, tokens[offsets[i]..offsets[i+1]) -> out[nseq][hidden] fp32.
 *      Encoder models (HB_ARCH_BERT): CLS row + L2.  Decoder models (HB_ARCH_LLAMA): last-token pooling of the final-norm
 *      hidden state + L2 — the `--task embed` mode of the reference's default embedding model, a Qwen2-style decoder
 *      (api/pkg/model/models.go:421-433); needs an engine that is not running generations concurrently. ---- */
int hb_embed(hb_engine* e, const int32_t* tokens, const int32_t* offsets, int32_t nseq, float* out);

/* ---- tokenizer (scope row F2): a HF tokenizer.json loaded natively — byte-level BPE with the Llama-3 pre-tokenizer
 *      (decoders) or BERT WordPiece with BertNormalizer / BertPreTokenizer (encoders such as bge) — bit-exact with
 *      `tokenizers` 0.22 on those pipelines; what the backends do inside their child process before the first kernel runs.
 *      encode/decode/chat return HB_ERR_BUSY with the needed size in *n / *len when the caller's buffer is too small. ---- */
typedef struct hb_tokenizer hb_tokenizer;
int hb_tok_load(const char* tokenizer_json_path, hb_tokenizer** out);
void hb_tok_free(hb_tokenizer* t);
int32_t hb_tok_vocab_size(hb_tokenizer* t);
int32_t hb_tok_token_id(hb_tokenizer* t, const char* token); /* -1 if absent; special tokens included */
/* parse_special: 0 = plain text; 1 = special-token strings inside the text become their ids; 2 = 1 + the model's framing
 * ([CLS] ... [SEP] for WordPiece encoders, <|begin_of_text|> for Llama-3 BPE): HF's encode(add_special_tokens=True) */
int hb_tok_encode(hb_tokenizer* t, const char* utf8, int32_t parse_special, int32_t* out, int32_t cap, int32_t* n);
int hb_tok_decode(hb_tokenizer* t, const int32_t* ids, int32_t n_ids, int32_t skip_special, char* out, size_t cap, size_t* len);
/* Llama-3 instruct template over (role, content) pairs, ending with the assistant header */
int hb_tok_chat_llama3(hb_tokenizer* t, const char* const* roles, const char* const* contents, int32_t n_msgs, int32_t* out,
                       int32_t cap, int32_t* n);

/* ---- response encoding (scope row F1 / a5): the embedding route answers with every vector as JSON decimal text
 *      (`data[i].embedding []float32`, api/pkg/runner/openai_embedding_handlers.go:110-772) and that encoding is where a
 *      scripting-language front spends its time (0.35 ms per 768-wide vector in CPython).  Writes `[v0,v1,...]` with the
 *      shortest text that parses back to the same float (what Go's encoding/json emits for float32); non-finite values
 *      become null.  Returns the number of bytes the array needs (no terminator); nothing is written beyond `cap`, so
 *      a return value > cap means "call again with a larger buffer".  Host-only: needs no GPU and no engine. ---- */
size_t hb_json_f32_array(const float* v, size_t n, char* out, size_t cap);

int hb_get_stats(hb_engine* e, hb_stats* out);
/* event-timed spans around every launch (measurement aid; disables CUDA-graph replay while on) */
int hb_set_profile(hb_engine* e, int32_t on);

#ifdef __cplusplus
}
#endif
#endif /* HELIX_B200_H_ */
