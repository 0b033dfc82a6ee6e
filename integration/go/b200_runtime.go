// Package runner — B200Runtime: the in-process Blackwell engine behind the runner.Runtime interface
// (api/pkg/runner/slot.go:46-57).  Drop this file, b200_front.go and b200_safetensors.go next to vllm_runtime.go;
// INTEGRATION.md has the lines that construct it in Slot.Create.
//
// This image has no Go toolchain, so these files are never compiled here.  What IS compiled and run is the same call
// sequence from plain C (tests/abi_host.c, driven on the GPU by tests/test_features_gpu.py) and the Python mirror
// helix_b200/runtime.py; every C call below appears in abi_host.c with the same arguments.
package runner

/*
#cgo CFLAGS: -I${SRCDIR}/../../../../helix-b200/include
#cgo LDFLAGS: -L${SRCDIR}/../../../../helix-b200/helix_b200 -lhelixb200 -lcudart
#include <stdlib.h>
#include "helix_b200.h"
*/
import "C"

import (
	"context"
	"errors"
	"fmt"
	"strconv"
	"strings"
	"sync"
	"unsafe"

	"github.com/helixml/helix/api/pkg/types"
)

type B200RuntimeParams struct {
	Model                  string
	GPUIndex               int      // CreateRunnerSlotAttributes.GPUIndex (types/runner.go:92-104)
	ModelMemoryRequirement uint64   // bytes the scheduler packed this slot with (scheduler/model_allocation.go:39-60)
	ContextLength          int64    // optional
	Args                   []string // vLLM-style args from the scheduler (scheduler/runner.go:1187-1259,1344-1397)
	Desc                   C.hb_model_desc
	CheckpointDir          string // HF checkpoint directory (*.safetensors) or a llama.cpp / Ollama *.gguf blob; empty = random init from Seed (benchmarks)
	Seed                   uint64
	Tokenizer              Tokenizer // see b200_front.go
	// Replica load (SURVEY.md §8e): when World > 1, rank 0 loads the checkpoint and every rank calls
	// hb_model_load_broadcast with the same 128-byte id (from NewReplicaID on the root).
	ReplicaID   []byte
	ReplicaRank int
	ReplicaSize int
}

type B200Runtime struct {
	p     B200RuntimeParams
	mu    sync.Mutex
	eng   *C.hb_engine
	front *openAIFront // serves /v1/chat/completions, /v1/embeddings, /v1/models on 127.0.0.1:<freeport>
	embed bool
	cfg   C.hb_engine_cfg
}

var _ Runtime = &B200Runtime{}

// GenParams are the request fields the runner forwards untouched to its backend
// (openai.ChatCompletionRequest, api/pkg/runner/openai_chat_handlers.go:100-175).
type GenParams struct {
	MaxTokens        int
	Temperature      float32
	TopP             float32 // 0 (absent) = nucleus filtering off
	TopK             int
	Seed             uint64
	EOS              int32 // < 0: none
	LogProbs         int   // 0 off; n >= 1: chosen token + n-1 alternatives per generated token
	PresencePenalty  float32
	FrequencyPenalty float32
}

// NewReplicaID is ncclGetUniqueId through the library: the Go host carries no NCCL binding of its own.
func NewReplicaID() ([]byte, error) {
	id := make([]byte, C.HB_REPLICA_ID_BYTES)
	if rc := C.hb_replica_unique_id(unsafe.Pointer(&id[0])); rc != C.HB_OK {
		return nil, fmt.Errorf("helix-b200: hb_replica_unique_id: %d (libnccl.so.2 missing?)", int(rc))
	}
	return id, nil
}

func NewB200Runtime(_ context.Context, p B200RuntimeParams) (*B200Runtime, error) {
	r := &B200Runtime{p: p}
	r.cfg.device = C.int32_t(p.GPUIndex)
	r.cfg.memory_budget_bytes = C.uint64_t(p.ModelMemoryRequirement)
	r.cfg.max_seqs = 256 // types/memory.go:11
	r.cfg.kv_page_size = 64
	r.cfg.use_cuda_graphs = 1
	r.cfg.enable_prefix_cache = 1 // vLLM V1's default: Helix sessions resend the whole conversation every turn
	r.cfg.decode_with_prefill = 1 // running streams keep decoding while a long prompt is prefilled (mixed steps)
	if p.ContextLength > 0 {
		r.cfg.max_ctx = C.int32_t(p.ContextLength)
	}
	for i := 0; i+1 < len(p.Args); i++ {
		switch p.Args[i] {
		case "--max-num-seqs":
			if v, err := strconv.Atoi(p.Args[i+1]); err == nil {
				r.cfg.max_seqs = C.int32_t(v)
			}
		case "--max-model-len":
			if v, err := strconv.Atoi(p.Args[i+1]); err == nil {
				r.cfg.max_ctx = C.int32_t(v)
			}
		case "--max-num-batched-tokens":
			if v, err := strconv.Atoi(p.Args[i+1]); err == nil {
				r.cfg.max_batched_tokens = C.int32_t(v) // prefill step budget; longer prompts are chunked
			}
		case "--task":
			r.embed = p.Args[i+1] == "embed"
		}
	}
	for _, a := range p.Args {
		if a == "--no-enable-prefix-caching" {
			r.cfg.enable_prefix_cache = 0
		}
	}
	if r.embed {
		r.cfg.enable_prefix_cache = 0
	}
	return r, nil
}

func (r *B200Runtime) lastError() error {
	return fmt.Errorf("helix-b200: %s", C.GoString(C.hb_last_error(r.eng)))
}

// loadWeights: checkpoint -> hb_model_load_begin / hb_model_tensor_set / hb_model_load_finish (bf16 rows by their HF
// names), or random init for benchmarks; replicas other than rank 0 receive rank 0's arena by one NCCL broadcast.
func (r *B200Runtime) loadWeights() error {
	world := r.p.ReplicaSize
	if world > 1 && r.p.ReplicaRank != 0 {
		var sec C.double
		if rc := C.hb_model_load_broadcast(r.eng, &r.p.Desc, unsafe.Pointer(&r.p.ReplicaID[0]), C.int32_t(r.p.ReplicaRank),
			C.int32_t(world), &sec); rc != C.HB_OK {
			return r.lastError()
		}
		return nil
	}
	if r.p.CheckpointDir == "" {
		if rc := C.hb_model_load_random(r.eng, &r.p.Desc, C.uint64_t(r.p.Seed)); rc != C.HB_OK {
			return r.lastError()
		}
	} else if strings.HasSuffix(r.p.CheckpointDir, ".gguf") {
		// an Ollama blob (the reference's default catalogue format): description from the file's metadata, tensors
		// dequantised to bf16 inside the library
		cpath := C.CString(r.p.CheckpointDir)
		defer C.free(unsafe.Pointer(cpath))
		if rc := C.hb_gguf_describe(cpath, &r.p.Desc); rc != C.HB_OK {
			return fmt.Errorf("helix-b200: %s", C.GoString(C.hb_last_error(nil)))
		}
		if rc := C.hb_model_load_gguf(r.eng, cpath); rc != C.HB_OK {
			return r.lastError()
		}
	} else {
		if rc := C.hb_model_load_begin(r.eng, &r.p.Desc); rc != C.HB_OK {
			return r.lastError()
		}
		err := forEachSafetensor(r.p.CheckpointDir, func(name string, bf16 []uint16) error {
			cname := C.CString(name)
			defer C.free(unsafe.Pointer(cname))
			rc := C.hb_model_tensor_set(r.eng, cname, unsafe.Pointer(&bf16[0]), C.size_t(len(bf16)))
			if rc == C.HB_ERR_NOT_FOUND { // tensors the engine does not use (e.g. rotary_emb.inv_freq buffers)
				return nil
			}
			if rc != C.HB_OK {
				return r.lastError()
			}
			return nil
		})
		if err != nil {
			return err
		}
		if rc := C.hb_model_load_finish(r.eng); rc != C.HB_OK { // fails if any tensor of the description is missing
			return r.lastError()
		}
	}
	if world > 1 { // rank 0: send the loaded arena to the other replicas
		var sec C.double
		if rc := C.hb_model_load_broadcast(r.eng, &r.p.Desc, unsafe.Pointer(&r.p.ReplicaID[0]), 0, C.int32_t(world), &sec); rc != C.HB_OK {
			return r.lastError()
		}
	}
	return nil
}

// Start: engine on gpu_index inside the slot's budget, weights, step loop, HTTP front.  On any failure nothing stays
// allocated (Slot.Create's deferred Stop would also clean up, slot.go:113-140).
func (r *B200Runtime) Start(ctx context.Context) error {
	r.mu.Lock()
	defer r.mu.Unlock()
	if rc := C.hb_engine_create(&r.cfg, &r.eng); rc != C.HB_OK {
		return fmt.Errorf("helix-b200: create: %s", C.GoString(C.hb_last_error(nil)))
	}
	fail := func(err error) error {
		C.hb_engine_destroy(r.eng)
		r.eng = nil
		return err
	}
	if err := r.loadWeights(); err != nil {
		return fail(err)
	}
	if r.p.Desc.arch == C.HB_ARCH_LLAMA {
		if rc := C.hb_engine_start(r.eng); rc != C.HB_OK {
			return fail(r.lastError())
		}
	}
	front, err := newOpenAIFront(r) // net/http handlers calling Generate / hb_embed below
	if err != nil {
		return fail(err)
	}
	r.front = front
	return nil
}

// Stop releases ALL device memory synchronously (slot.go:113-140; server.go:801-817 then polls nvidia-smi).
// The front is closed first (no new handlers); hb_engine_destroy cancels open requests and wakes every goroutine still
// parked in hb_wait before it frees anything.
func (r *B200Runtime) Stop() error {
	r.mu.Lock()
	defer r.mu.Unlock()
	if r.front != nil {
		r.front.Close()
		r.front = nil
	}
	if r.eng != nil {
		C.hb_engine_destroy(r.eng)
		r.eng = nil
	}
	return nil
}

func (r *B200Runtime) PullModel(_ context.Context, _ string, progress func(PullProgress) error) error {
	return progress(PullProgress{Status: "success", Completed: 1, Total: 1})
}

// Warm (vllm_runtime.go:391-489 sends "Say the word 'warm'."): one short generation / encode.
func (r *B200Runtime) Warm(ctx context.Context, _ string) error {
	if r.embed {
		_, err := r.Embed([][]int32{{1, 2, 3}})
		return err
	}
	_, err := r.Generate(ctx, []int32{1, 2, 3, 4}, GenParams{MaxTokens: 2, EOS: -1}, func([]int32, []TokenLogProbs) error { return nil })
	return err
}

func (r *B200Runtime) ListModels(context.Context) ([]string, error) { return []string{r.p.Model}, nil }
func (r *B200Runtime) Version() string {
	return "helix-b200/0.2 (abi " + strconv.Itoa(int(C.hb_abi_version())) + ")"
}
func (r *B200Runtime) Runtime() types.Runtime { return types.RuntimeVLLM } // Option A, SURVEY.md §8b
func (r *B200Runtime) URL() string            { return r.front.URL() }
func (r *B200Runtime) CommandLine() string {
	return "helix-b200 (in-process) " + strings.Join(r.p.Args, " ")
}

// Status: non-empty == running (scheduler/scheduler.go:940); a sticky CUDA error reports "".
func (r *B200Runtime) Status(context.Context) string {
	var st C.hb_stats
	if r.eng == nil || C.hb_get_stats(r.eng, &st) != C.HB_OK || st.cuda_error != 0 {
		return ""
	}
	return fmt.Sprintf("running kv_pages_free=%d/%d running=%d waiting=%d", int(st.kv_pages_free), int(st.kv_pages_total),
		int(st.running), int(st.waiting))
}

// TokenLogProbs is one row of hb_logprobs: IDs[0] is the sampled token, the rest the most likely tokens (descending).
type TokenLogProbs struct {
	IDs      []int32
	LogProbs []float32
}

var errEngineAborted = errors.New("helix-b200: generation aborted by the engine")

// Generate streams the token ids of one request; emit is called once per poll (one SSE chunk each).  finished: 1 = the
// sequence ended normally (max_tokens / EOS), 2 = cancelled or failed.  The request record is released only after the
// step loop has retired the sequence (releasing a RUNNING request is an error and would leak the record).
func (r *B200Runtime) Generate(ctx context.Context, prompt []int32, gp GenParams, emit func([]int32, []TokenLogProbs) error) (finished int, err error) {
	if len(prompt) == 0 {
		return 0, errors.New("helix-b200: empty prompt")
	}
	var sp C.hb_sampling // zero value = greedy, everything optional off
	sp.temperature = C.float(gp.Temperature)
	sp.seed = C.uint64_t(gp.Seed)
	sp.max_tokens = C.int32_t(gp.MaxTokens)
	sp.eos_token = C.int32_t(gp.EOS)
	sp.top_k = C.int32_t(gp.TopK)
	sp.top_p = C.float(gp.TopP)
	sp.logprobs = C.int32_t(gp.LogProbs)
	sp.presence_penalty = C.float(gp.PresencePenalty)
	sp.frequency_penalty = C.float(gp.FrequencyPenalty)
	var id C.uint64_t
	if rc := C.hb_submit(r.eng, (*C.int32_t)(unsafe.Pointer(&prompt[0])), C.int32_t(len(prompt)), &sp, &id); rc != C.HB_OK {
		return 0, r.lastError()
	}
	buf := make([]int32, 256)
	lpIDs := make([]int32, 256*C.HB_MAX_LOGPROBS)
	lpVals := make([]float32, 256*C.HB_MAX_LOGPROBS)
	lpRow := 0
	// retire: cancel if still active, wait until the step loop has let go of the sequence, then drop the record
	retire := func(active bool) {
		if active {
			C.hb_cancel(r.eng, id)
			for i := 0; i < 1000; i++ {
				var n, fin C.int32_t
				if C.hb_poll(r.eng, id, (*C.int32_t)(unsafe.Pointer(&buf[0])), C.int32_t(len(buf)), &n, &fin) != C.HB_OK || fin != 0 {
					break
				}
				C.hb_wait(r.eng, id, 10)
			}
		}
		C.hb_release(r.eng, id)
	}
	for {
		if ctx.Err() != nil { // client went away: frees the sequence's KV pages at the next step boundary
			retire(true)
			return 2, ctx.Err()
		}
		C.hb_wait(r.eng, id, 100)
		var n, fin C.int32_t
		if rc := C.hb_poll(r.eng, id, (*C.int32_t)(unsafe.Pointer(&buf[0])), C.int32_t(len(buf)), &n, &fin); rc != C.HB_OK {
			err := r.lastError()
			retire(true)
			return 0, err
		}
		if n > 0 {
			var lps []TokenLogProbs
			if gp.LogProbs > 0 {
				var rows, width C.int32_t
				if rc := C.hb_logprobs(r.eng, id, C.int32_t(lpRow), n, (*C.int32_t)(unsafe.Pointer(&lpIDs[0])),
					(*C.float)(unsafe.Pointer(&lpVals[0])), &rows, &width); rc == C.HB_OK {
					w := int(width)
					for i := 0; i < int(rows); i++ {
						lps = append(lps, TokenLogProbs{IDs: append([]int32(nil), lpIDs[i*w:(i+1)*w]...),
							LogProbs: append([]float32(nil), lpVals[i*w:(i+1)*w]...)})
					}
					lpRow += int(rows)
				}
			}
			if err := emit(buf[:int(n)], lps); err != nil {
				retire(fin == 0)
				return 2, err
			}
		}
		if fin != 0 {
			retire(false)
			if fin == 2 {
				return 2, errEngineAborted // FAILED / CANCELLED by the engine: not a normal completion
			}
			return 1, nil
		}
	}
}

// Embed: one hb_embed call for a batch of token sequences -> L2-normalised vectors (CLS pooling for encoders, last-token
// pooling for decoder embedders).
func (r *B200Runtime) Embed(seqs [][]int32) ([][]float32, error) {
	if len(seqs) == 0 {
		return nil, nil
	}
	var toks []int32
	offs := []int32{0}
	for _, s := range seqs {
		if len(s) == 0 {
			return nil, errors.New("helix-b200: empty sequence")
		}
		toks = append(toks, s...)
		offs = append(offs, int32(len(toks)))
	}
	hidden := int(r.p.Desc.hidden)
	out := make([]float32, len(seqs)*hidden)
	if rc := C.hb_embed(r.eng, (*C.int32_t)(unsafe.Pointer(&toks[0])), (*C.int32_t)(unsafe.Pointer(&offs[0])), C.int32_t(len(seqs)),
		(*C.float)(unsafe.Pointer(&out[0]))); rc != C.HB_OK {
		return nil, r.lastError()
	}
	vecs := make([][]float32, len(seqs))
	for i := range seqs {
		vecs[i] = out[i*hidden : (i+1)*hidden]
	}
	return vecs, nil
}
