// Package runner — openAIFront: what B200Runtime.URL() serves.  The reference's handlers build a go-openai client on
// slot.URL()+"/v1" (openai_chat_handlers.go:100, openai_embedding_handlers.go:291, openai_model_handlers.go:36), so this
// front speaks exactly that dialect: POST /v1/chat/completions (JSON, or SSE when stream=true: one `data: {chunk}`
// per poll, a final chunk with a non-empty finish_reason, then `data: [DONE]`), POST /v1/embeddings, GET /v1/models.
// Source only (no Go toolchain in this image); helix_b200/server.py is the executable mirror the tests drive.
package runner

/*
#include "helix_b200.h"
*/
import "C"

import (
	"encoding/json"
	"fmt"
	"net"
	"net/http"
	"time"

	openai "github.com/sashabaranov/go-openai"
)

// Tokenizer is supplied by the runner (tiktoken-go / daulet/tokenizers are already in go.mod:76,128).
type Tokenizer interface {
	EncodeChat(messages []openai.ChatCompletionMessage) []int32
	Encode(text string) []int32
	Decode(ids []int32) string
	EOS() int32
}

type openAIFront struct {
	rt  *B200Runtime
	tok Tokenizer
	ln  net.Listener
	srv *http.Server
}

func newOpenAIFront(rt *B200Runtime) (*openAIFront, error) {
	ln, err := net.Listen("tcp", "127.0.0.1:0") // free port, like freeport.GetFreePort() for the child processes
	if err != nil {
		return nil, err
	}
	f := &openAIFront{rt: rt, tok: rt.p.Tokenizer, ln: ln}
	mux := http.NewServeMux()
	mux.HandleFunc("/v1/models", f.models)
	mux.HandleFunc("/v1/chat/completions", f.chat)
	mux.HandleFunc("/v1/embeddings", f.embeddings)
	f.srv = &http.Server{Handler: mux}
	go f.srv.Serve(ln) //nolint:errcheck
	return f, nil
}

func (f *openAIFront) URL() string { return "http://" + f.ln.Addr().String() }
func (f *openAIFront) Close()      { _ = f.srv.Close() }

func (f *openAIFront) models(w http.ResponseWriter, _ *http.Request) {
	_ = json.NewEncoder(w).Encode(openai.ModelsList{Models: []openai.Model{{ID: f.rt.p.Model, Object: "model", OwnedBy: "helix-b200"}}})
}

func (f *openAIFront) chat(w http.ResponseWriter, r *http.Request) {
	var req openai.ChatCompletionRequest
	if err := json.NewDecoder(http.MaxBytesReader(w, r.Body, 10*1024*1024)).Decode(&req); err != nil { // openai_chat_handlers.go:40
		http.Error(w, err.Error(), http.StatusBadRequest)
		return
	}
	if req.Model != "" && req.Model != f.rt.p.Model {
		http.Error(w, fmt.Sprintf("model mismatch, expecting %s", f.rt.p.Model), http.StatusBadRequest)
		return
	}
	prompt := f.tok.EncodeChat(req.Messages)
	maxTokens := req.MaxTokens
	if maxTokens == 0 {
		maxTokens = 256
	}
	id, created := "chatcmpl-"+randomID(), time.Now().Unix()
	chunk := func(delta openai.ChatCompletionStreamChoiceDelta, finish openai.FinishReason) openai.ChatCompletionStreamResponse {
		return openai.ChatCompletionStreamResponse{ID: id, Object: "chat.completion.chunk", Created: created, Model: f.rt.p.Model,
			Choices: []openai.ChatCompletionStreamChoice{{Index: 0, Delta: delta, FinishReason: finish}}}
	}
	var full string
	n := 0
	var all []int32 // every generated id so far: the text is decoded from the whole sequence (see emitStable)
	emitted := 0
	var flusher http.Flusher
	if req.Stream {
		w.Header().Set("Content-Type", "text/event-stream")
		w.Header().Set("Cache-Control", "no-cache")
		flusher, _ = w.(http.Flusher)
		writeSSE(w, flusher, chunk(openai.ChatCompletionStreamChoiceDelta{Role: "assistant"}, ""))
	}
	fin, err := f.rt.Generate(r.Context(), prompt, maxTokens, req.Temperature, req.TopP, uint64(derefInt(req.Seed)), func(ids []int32) error {
		n += len(ids)
		all = append(all, dropToken(ids, f.tok.EOS())...)
		// A character whose bytes / pieces straddle two polls must come out whole: decode everything, release only the
		// new stable suffix, hold back a trailing U+FFFD (an incomplete sequence so far).  Mirrors server.py StreamDecoder.
		text := emitStable(f.tok.Decode(all), &emitted, false)
		full += text
		if req.Stream && text != "" {
			return writeSSE(w, flusher, chunk(openai.ChatCompletionStreamChoiceDelta{Content: text}, ""))
		}
		return nil
	})
	if err != nil && fin == 0 {
		http.Error(w, err.Error(), http.StatusInternalServerError)
		return
	}
	if tail := emitStable(f.tok.Decode(all), &emitted, true); tail != "" { // whatever was still held back
		full += tail
		if req.Stream {
			writeSSE(w, flusher, chunk(openai.ChatCompletionStreamChoiceDelta{Content: tail}, ""))
		}
	}
	reason := openai.FinishReasonStop
	if n >= maxTokens {
		reason = openai.FinishReasonLength
	}
	if req.Stream {
		writeSSE(w, flusher, chunk(openai.ChatCompletionStreamChoiceDelta{}, reason)) // closes the control-plane stream (helix_openai_client.go:197)
		fmt.Fprint(w, "data: [DONE]\n\n")
		return
	}
	_ = json.NewEncoder(w).Encode(openai.ChatCompletionResponse{ID: id, Object: "chat.completion", Created: created, Model: f.rt.p.Model,
		Choices: []openai.ChatCompletionChoice{{Index: 0, Message: openai.ChatCompletionMessage{Role: "assistant", Content: full}, FinishReason: reason}},
		Usage:   openai.Usage{PromptTokens: len(prompt), CompletionTokens: n, TotalTokens: len(prompt) + n}})
}

func (f *openAIFront) embeddings(w http.ResponseWriter, r *http.Request) {
	var req struct {
		Input json.RawMessage `json:"input"` // string | []string | [][]int (types/types.go:2707-2730)
		Model string          `json:"model"`
	}
	if err := json.NewDecoder(r.Body).Decode(&req); err != nil {
		http.Error(w, err.Error(), http.StatusBadRequest)
		return
	}
	seqs, err := decodeEmbeddingInput(req.Input, f.tok)
	if err != nil {
		http.Error(w, err.Error(), http.StatusBadRequest)
		return
	}
	// flatten -> one hb_embed call (a micro-batcher in front of this coalesces the RAG caller's 1-chunk requests)
	var toks []int32
	offs := []int32{0}
	for _, s := range seqs {
		toks = append(toks, s...)
		offs = append(offs, int32(len(toks)))
	}
	hidden := int(f.rt.p.Desc.hidden)
	out := make([]float32, len(seqs)*hidden)
	if rc := C.hb_embed(f.rt.eng, (*C.int32_t)(&toks[0]), (*C.int32_t)(&offs[0]), C.int32_t(len(seqs)), (*C.float)(&out[0])); rc != C.HB_OK {
		http.Error(w, f.rt.lastError().Error(), http.StatusInternalServerError)
		return
	}
	resp := openai.EmbeddingResponse{Object: "list", Model: openai.EmbeddingModel(f.rt.p.Model), Usage: openai.Usage{PromptTokens: len(toks), TotalTokens: len(toks)}}
	for i := range seqs {
		resp.Data = append(resp.Data, openai.Embedding{Object: "embedding", Index: i, Embedding: out[i*hidden : (i+1)*hidden]})
	}
	_ = json.NewEncoder(w).Encode(resp)
}

// emitStable returns the part of `decoded` not yet released; unless `final`, one trailing replacement rune stays held.
func emitStable(decoded string, emitted *int, final bool) string {
	r := []rune(decoded)
	stable := len(r)
	if !final && stable > 0 && r[stable-1] == '\uFFFD' {
		stable--
	}
	if stable <= *emitted {
		return ""
	}
	out := string(r[*emitted:stable])
	*emitted = stable
	return out
}

func writeSSE(w http.ResponseWriter, fl http.Flusher, v any) error {
	b, err := json.Marshal(v)
	if err != nil {
		return err
	}
	if _, err := fmt.Fprintf(w, "data: %s\n\n", b); err != nil {
		return err
	}
	if fl != nil {
		fl.Flush()
	}
	return nil
}
