// Package runner — B200Runtime: the in-process Blackwell engine behind the runner.Runtime interface
// (api/pkg/runner/slot.go:46-57).  Drop this file next to vllm_runtime.go; see INTEGRATION.md for the three
// lines that construct it in Slot.Create.  NOTE: this image has no Go toolchain, so this file is shipped as
// source only (never compiled here); helix_b200/runtime.py is its executable mirror and is what the tests drive.
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
	Seed                   uint64
	Tokenizer              Tokenizer // see b200_front.go
}

type B200Runtime struct {
	p      B200RuntimeParams
	mu     sync.Mutex
	eng    *C.hb_engine
	front  *openAIFront // serves /v1/chat/completions, /v1/embeddings, /v1/models on 127.0.0.1:<freeport>
	embed  bool
	cfg    C.hb_engine_cfg
}

var _ Runtime = &B200Runtime{}

func NewB200Runtime(_ context.Context, p B200RuntimeParams) (*B200Runtime, error) {
	r := &B200Runtime{p: p}
	r.cfg.device = C.int32_t(p.GPUIndex)
	r.cfg.memory_budget_bytes = C.uint64_t(p.ModelMemoryRequirement)
	r.cfg.max_seqs = 256 // types/memory.go:11
	r.cfg.kv_page_size = 64
	r.cfg.use_cuda_graphs = 1
	r.cfg.enable_prefix_cache = 1 // vLLM V1's default: Helix sessions resend the whole conversation every turn
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

// Start: engine on gpu_index inside the slot's budget, weights, step loop, HTTP front.
func (r *B200Runtime) Start(ctx context.Context) error {
	r.mu.Lock()
	defer r.mu.Unlock()
	if rc := C.hb_engine_create(&r.cfg, &r.eng); rc != C.HB_OK {
		return fmt.Errorf("helix-b200: create: %s", C.GoString(C.hb_last_error(nil)))
	}
	if rc := C.hb_model_load_random(r.eng, &r.p.Desc, C.uint64_t(r.p.Seed)); rc != C.HB_OK { // or load_begin/tensor_set/finish from safetensors
		err := r.lastError()
		C.hb_engine_destroy(r.eng)
		r.eng = nil
		return err
	}
	if r.p.Desc.arch == C.HB_ARCH_LLAMA {
		if rc := C.hb_engine_start(r.eng); rc != C.HB_OK {
			return r.lastError()
		}
	}
	front, err := newOpenAIFront(r) // net/http handlers calling Submit/Poll/Embed below
	if err != nil {
		return err
	}
	r.front = front
	return nil
}

// Stop releases ALL device memory synchronously (slot.go:113-140; server.go:801-817 then polls nvidia-smi).
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

func (r *B200Runtime) Warm(ctx context.Context, _ string) error {
	if r.embed {
		toks, offs := []int32{1, 2, 3}, []int32{0, 3}
		out := make([]float32, int(r.p.Desc.hidden))
		if rc := C.hb_embed(r.eng, (*C.int32_t)(&toks[0]), (*C.int32_t)(&offs[0]), 1, (*C.float)(&out[0])); rc != C.HB_OK {
			return r.lastError()
		}
		return nil
	}
	_, err := r.Generate(ctx, []int32{1, 2, 3, 4}, 2, 0, 0, func([]int32) error { return nil })
	return err
}

func (r *B200Runtime) ListModels(context.Context) ([]string, error) { return []string{r.p.Model}, nil }
func (r *B200Runtime) Version() string                              { return "helix-b200/0.1 (abi " + strconv.Itoa(int(C.hb_abi_version())) + ")" }
func (r *B200Runtime) Runtime() types.Runtime                       { return types.RuntimeVLLM } // Option A, SURVEY.md §8b
func (r *B200Runtime) URL() string                                  { return r.front.URL() }
func (r *B200Runtime) CommandLine() string                          { return "helix-b200 (in-process) " + strings.Join(r.p.Args, " ") }

// Status: non-empty == running (scheduler/scheduler.go:940); a sticky CUDA error reports "".
func (r *B200Runtime) Status(context.Context) string {
	var st C.hb_stats
	if r.eng == nil || C.hb_get_stats(r.eng, &st) != C.HB_OK || st.cuda_error != 0 {
		return ""
	}
	return fmt.Sprintf("running kv_pages_free=%d/%d running=%d waiting=%d", st.kv_pages_free, st.kv_pages_total, st.running, st.waiting)
}

// Generate streams token ids of one request; emit is called once per poll (one SSE chunk each).
func (r *B200Runtime) Generate(ctx context.Context, prompt []int32, maxTokens int, temperature, topP float32, seed uint64, emit func([]int32) error) (finished int, err error) {
	sp := C.hb_sampling{temperature: C.float(temperature), seed: C.uint64_t(seed), max_tokens: C.int32_t(maxTokens), eos_token: -1,
		top_p: C.float(topP)} // 0 (absent from the request) = nucleus filtering off
	var id C.uint64_t
	if rc := C.hb_submit(r.eng, (*C.int32_t)(unsafe.Pointer(&prompt[0])), C.int32_t(len(prompt)), &sp, &id); rc != C.HB_OK {
		return 0, r.lastError()
	}
	defer C.hb_release(r.eng, id)
	buf := make([]int32, 256)
	for {
		if ctx.Err() != nil {
			C.hb_cancel(r.eng, id) // frees the sequence's KV pages at the next step boundary
			return 2, ctx.Err()
		}
		C.hb_wait(r.eng, id, 100)
		var n, fin C.int32_t
		if rc := C.hb_poll(r.eng, id, (*C.int32_t)(&buf[0]), C.int32_t(len(buf)), &n, &fin); rc != C.HB_OK {
			return 0, r.lastError()
		}
		if n > 0 {
			if err := emit(buf[:n]); err != nil {
				C.hb_cancel(r.eng, id)
				return 2, err
			}
		}
		if fin != 0 {
			return int(fin), nil
		}
	}
}
