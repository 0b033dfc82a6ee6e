// Package runner — openAIFront: what B200Runtime.URL() serves.  The reference's handlers build a go-openai client on
// slot.URL()+"/v1" (openai_chat_handlers.go:100, openai_embedding_handlers.go:291, openai_model_handlers.go:36), so this
// front speaks exactly that dialect: POST /v1/chat/completions (JSON, or SSE when stream=true: `data: {chunk}` events, a
// final chunk per choice with a non-empty finish_reason, then `data: [DONE]`), POST /v1/embeddings, GET /v1/models.
// Pure Go over B200Runtime.Generate / Embed (no cgo in this file).  Source only (no Go toolchain in this image);
// helix_b200/server.py is the executable mirror the tests drive over real sockets (tests/test_server_cpu.py).
package runner

import (
	"crypto/rand"
	"encoding/hex"
	"encoding/json"
	"errors"
	"fmt"
	"net"
	"net/http"
	"strings"
	"sync"
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
	rt    *B200Runtime
	tok   Tokenizer
	ln    net.Listener
	srv   *http.Server
	embed *embedBatcher
}

// embedBatcher coalesces concurrent embedding requests into few hb_embed calls (scope row F4).  The reference's indexer
// keeps a fixed pool of workers that each send ONE chunk per request and the next one when the answer arrives
// (rag/rag_pgvector.go:70-83): encoding them one by one leaves the GPU idle.  A batch takes everything that is queued,
// then waits — embedWindow at most — until as many requests are in as the previous batch held: once the whole pool has
// reported in, nobody else is about to show up.  Mirrors helix_b200/server.py EmbedBatcher (measured there: 10 workers,
// 512-token chunks: 2.2k chunks/s at 4.3 ms per request; one worker alone 1.2 ms).
type embedBatcher struct {
	rt   *B200Runtime
	jobs chan *embedJob
	quit chan struct{}
}

type embedJob struct {
	seqs [][]int32
	done chan embedResult
}

type embedResult struct {
	vecs [][]float32
	err  error
}

const (
	embedWindow  = 2 * time.Millisecond
	embedMaxSeqs = 256
)

func newEmbedBatcher(rt *B200Runtime) *embedBatcher {
	b := &embedBatcher{rt: rt, jobs: make(chan *embedJob, 4096), quit: make(chan struct{})}
	go b.run()
	return b
}

func (b *embedBatcher) Embed(seqs [][]int32) ([][]float32, error) {
	j := &embedJob{seqs: seqs, done: make(chan embedResult, 1)}
	select {
	case b.jobs <- j:
	case <-b.quit:
		return nil, errors.New("helix-b200: runtime stopped")
	}
	select {
	case r := <-j.done:
		return r.vecs, r.err
	case <-b.quit:
		return nil, errors.New("helix-b200: runtime stopped")
	}
}

func (b *embedBatcher) Close() { close(b.quit) }

func (b *embedBatcher) run() {
	expect := 1 // requests the next batch waits for: the size of the previous one
	for {
		var batch []*embedJob
		select {
		case j := <-b.jobs:
			batch = append(batch, j)
		case <-b.quit:
			return
		}
		nseq := len(batch[0].seqs)
		timer := time.NewTimer(embedWindow)
	collect:
		for nseq < embedMaxSeqs {
			if len(batch) < expect {
				select { // wait for company, but not past the window
				case j := <-b.jobs:
					batch = append(batch, j)
					nseq += len(j.seqs)
				case <-timer.C:
					break collect
				case <-b.quit:
					timer.Stop()
					return
				}
			} else {
				select { // whatever else is already queued rides along
				case j := <-b.jobs:
					batch = append(batch, j)
					nseq += len(j.seqs)
				default:
					break collect
				}
			}
		}
		timer.Stop()
		expect = len(batch)
		flat := make([][]int32, 0, nseq)
		for _, j := range batch {
			flat = append(flat, j.seqs...)
		}
		vecs, err := b.rt.Embed(flat)
		if err != nil && len(batch) > 1 {
			for _, j := range batch { // one bad sequence must not poison its neighbours: answer each request alone
				v, e := b.rt.Embed(j.seqs)
				j.done <- embedResult{vecs: v, err: e}
			}
			continue
		}
		k := 0
		for _, j := range batch {
			if err != nil {
				j.done <- embedResult{err: err}
				continue
			}
			j.done <- embedResult{vecs: vecs[k : k+len(j.seqs)]}
			k += len(j.seqs)
		}
	}
}

func newOpenAIFront(rt *B200Runtime) (*openAIFront, error) {
	ln, err := net.Listen("tcp", "127.0.0.1:0") // free port, like freeport.GetFreePort() for the child processes
	if err != nil {
		return nil, err
	}
	f := &openAIFront{rt: rt, tok: rt.p.Tokenizer, ln: ln, embed: newEmbedBatcher(rt)}
	mux := http.NewServeMux()
	mux.HandleFunc("/v1/models", f.models)
	mux.HandleFunc("/v1/chat/completions", f.chat)
	mux.HandleFunc("/v1/embeddings", f.embeddings)
	f.srv = &http.Server{Handler: mux}
	go f.srv.Serve(ln) //nolint:errcheck
	return f, nil
}

func (f *openAIFront) URL() string { return "http://" + f.ln.Addr().String() }

// Close stops accepting connections and closes the open ones; handlers notice through their request context and retire
// their sequences (Generate cancels + releases) before B200Runtime.Stop destroys the engine.
func (f *openAIFront) Close() {
	_ = f.srv.Close()
	f.embed.Close()
}

func randomID() string {
	var b [12]byte
	_, _ = rand.Read(b[:])
	return hex.EncodeToString(b[:])
}

func derefInt(p *int) int {
	if p == nil {
		return 0
	}
	return *p
}

// dropToken returns ids without any occurrence of tok (the EOS id never reaches the visible text).
func dropToken(ids []int32, tok int32) []int32 {
	out := make([]int32, 0, len(ids))
	for _, t := range ids {
		if t != tok {
			out = append(out, t)
		}
	}
	return out
}

// decodeEmbeddingInput accepts the three forms the reference forwards (types/types.go:2707-2730):
// "text" | ["text", ...] | [[id, ...], ...]; a flat [id, ...] is one sequence.
func decodeEmbeddingInput(raw json.RawMessage, tok Tokenizer) ([][]int32, error) {
	var s string
	if json.Unmarshal(raw, &s) == nil {
		return [][]int32{tok.Encode(s)}, nil
	}
	var ss []string
	if json.Unmarshal(raw, &ss) == nil && len(ss) > 0 {
		out := make([][]int32, len(ss))
		for i, x := range ss {
			out[i] = tok.Encode(x)
		}
		return out, nil
	}
	var flat []int32
	if json.Unmarshal(raw, &flat) == nil && len(flat) > 0 {
		return [][]int32{flat}, nil
	}
	var nested [][]int32
	if json.Unmarshal(raw, &nested) == nil && len(nested) > 0 {
		return nested, nil
	}
	return nil, errors.New("input must be a string, a list of strings or token arrays")
}

func (f *openAIFront) models(w http.ResponseWriter, _ *http.Request) {
	_ = json.NewEncoder(w).Encode(openai.ModelsList{Models: []openai.Model{{ID: f.rt.p.Model, Object: "model", OwnedBy: "helix-b200"}}})
}

func jsonError(w http.ResponseWriter, code int, kind, msg string) {
	w.Header().Set("Content-Type", "application/json")
	w.WriteHeader(code)
	_ = json.NewEncoder(w).Encode(map[string]any{"error": map[string]string{"message": msg, "type": kind}})
}

// firstStop returns the index of the earliest stop string in text, or -1.
func firstStop(text string, stops []string) int {
	cut := -1
	for _, s := range stops {
		if s == "" {
			continue
		}
		if i := strings.Index(text, s); i >= 0 && (cut < 0 || i < cut) {
			cut = i
		}
	}
	return cut
}

func (f *openAIFront) chat(w http.ResponseWriter, r *http.Request) {
	var req openai.ChatCompletionRequest
	if err := json.NewDecoder(http.MaxBytesReader(w, r.Body, 10*1024*1024)).Decode(&req); err != nil { // openai_chat_handlers.go:40
		jsonError(w, http.StatusBadRequest, "invalid_request_error", err.Error())
		return
	}
	if req.Model != "" && req.Model != f.rt.p.Model {
		jsonError(w, http.StatusBadRequest, "invalid_request_error", fmt.Sprintf("model mismatch, expecting %s", f.rt.p.Model))
		return
	}
	nChoices := req.N
	if nChoices == 0 {
		nChoices = 1
	}
	if nChoices < 1 || nChoices > 16 || req.TopLogProbs < 0 || req.TopLogProbs > 20 {
		jsonError(w, http.StatusBadRequest, "invalid_request_error", "n must be in [1,16], top_logprobs in [0,20]")
		return
	}
	prompt := f.tok.EncodeChat(req.Messages)
	if len(prompt) == 0 {
		jsonError(w, http.StatusBadRequest, "invalid_request_error", "empty prompt")
		return
	}
	maxTokens := req.MaxTokens
	if maxTokens == 0 {
		maxTokens = req.MaxCompletionTokens
	}
	if maxTokens == 0 {
		maxTokens = 256
	}
	gp := GenParams{MaxTokens: maxTokens, Temperature: req.Temperature, TopP: req.TopP, Seed: uint64(derefInt(req.Seed)), EOS: f.tok.EOS(),
		PresencePenalty: req.PresencePenalty, FrequencyPenalty: req.FrequencyPenalty}
	if req.LogProbs {
		gp.LogProbs = 1 + req.TopLogProbs
	}
	id, created := "chatcmpl-"+randomID(), time.Now().Unix()

	// One goroutine per choice (n > 1: same prompt, seed + i; the prefix cache shares the prompt's KV pages).  All writes
	// to the response go through `mu`; the headers of a stream go out with the first event, so a failure before any token
	// is still a real 4xx/5xx.
	type choiceState struct {
		text   string
		n      int
		reason openai.FinishReason
		lps    []openai.LogProb
		err    error
	}
	states := make([]choiceState, nChoices)
	var mu sync.Mutex
	headerSent := false
	var flusher http.Flusher
	sendSSE := func(v any) error { // mu held
		if !headerSent {
			w.Header().Set("Content-Type", "text/event-stream")
			w.Header().Set("Cache-Control", "no-cache")
			flusher, _ = w.(http.Flusher)
			headerSent = true
		}
		return writeSSE(w, flusher, v)
	}
	chunk := func(idx int, delta openai.ChatCompletionStreamChoiceDelta, finish openai.FinishReason) openai.ChatCompletionStreamResponse {
		return openai.ChatCompletionStreamResponse{ID: id, Object: "chat.completion.chunk", Created: created, Model: f.rt.p.Model,
			Choices: []openai.ChatCompletionStreamChoice{{Index: idx, Delta: delta, FinishReason: finish}}}
	}
	var wg sync.WaitGroup
	for ci := 0; ci < nChoices; ci++ {
		wg.Add(1)
		go func(ci int) {
			defer wg.Done()
			st := &states[ci]
			g := gp
			g.Seed += uint64(ci)
			dec := &streamDecoder{tok: f.tok} // ids -> text without re-decoding the whole generation at every poll
			pending := ""  // decoded text not yet released because it may be the beginning of a stop string
			stopped := false
			release := func(text string, final bool) (string, bool) { // -> text that may be shown now, stop hit?
				pending += text
				if cut := firstStop(pending, req.Stop); cut >= 0 {
					out := pending[:cut]
					pending = ""
					return out, true
				}
				if final {
					out := pending
					pending = ""
					return out, false
				}
				keep := 0 // longest suffix that is a proper prefix of some stop string
				for _, s := range req.Stop {
					for k := len(s) - 1; k > keep; k-- {
						if k <= len(pending) && strings.HasSuffix(pending, s[:k]) {
							keep = k
						}
					}
				}
				out := pending[:len(pending)-keep]
				pending = pending[len(pending)-keep:]
				return out, false
			}
			errStop := errors.New("stop sequence")
			started := false
			fin, err := f.rt.Generate(r.Context(), prompt, g, func(ids []int32, lps []TokenLogProbs) error {
				mu.Lock()
				defer mu.Unlock()
				if req.Stream && !started {
					started = true
					if err := sendSSE(chunk(ci, openai.ChatCompletionStreamChoiceDelta{Role: "assistant"}, "")); err != nil {
						return err
					}
				}
				st.n += len(ids)
				fresh := dropToken(ids, f.tok.EOS())
				for _, lp := range lps {
					e := openai.LogProb{Token: f.tok.Decode(lp.IDs[:1]), LogProb: float64(lp.LogProbs[0])}
					for k := 1; k < len(lp.IDs); k++ {
						if lp.IDs[k] >= 0 {
							e.TopLogProbs = append(e.TopLogProbs, openai.TopLogProbs{Token: f.tok.Decode(lp.IDs[k : k+1]), LogProb: float64(lp.LogProbs[k])})
						}
					}
					st.lps = append(st.lps, e)
				}
				// A character whose bytes / pieces straddle two polls must come out whole (streamDecoder holds an unfinished
				// one back); text that may still become a stop string is held by release().  Mirrors server.py.
				text, hit := release(dec.feed(fresh, false), false)
				st.text += text
				if req.Stream && text != "" {
					if err := sendSSE(chunk(ci, openai.ChatCompletionStreamChoiceDelta{Content: text}, "")); err != nil {
						return err
					}
				}
				if hit {
					stopped = true
					return errStop // Generate cancels the sequence (frees its KV pages) and releases the record
				}
				return nil
			})
			mu.Lock()
			defer mu.Unlock()
			if err != nil && !stopped {
				st.err = err
				return
			}
			if !stopped {
				tail, _ := release(dec.feed(nil, true), true)
				st.text += tail
				if req.Stream && tail != "" {
					_ = sendSSE(chunk(ci, openai.ChatCompletionStreamChoiceDelta{Content: tail}, ""))
				}
			}
			st.reason = openai.FinishReasonStop
			if !stopped && fin == 1 && st.n >= maxTokens {
				st.reason = openai.FinishReasonLength
			}
			if req.Stream {
				_ = sendSSE(chunk(ci, openai.ChatCompletionStreamChoiceDelta{}, st.reason)) // closes the control-plane stream (helix_openai_client.go:197)
			}
		}(ci)
	}
	wg.Wait()
	var firstErr error
	total := 0
	for i := range states {
		if states[i].err != nil && firstErr == nil {
			firstErr = states[i].err
		}
		total += states[i].n
	}
	if firstErr != nil {
		if headerSent { // the 200 is out: an SSE error event, no finish_reason / [DONE] — never "stop" on a truncated answer
			_ = writeSSE(w, flusher, map[string]any{"error": map[string]string{"message": firstErr.Error(), "type": "server_error"}})
		} else {
			jsonError(w, http.StatusInternalServerError, "server_error", firstErr.Error())
		}
		return
	}
	if req.Stream {
		fmt.Fprint(w, "data: [DONE]\n\n")
		return
	}
	resp := openai.ChatCompletionResponse{ID: id, Object: "chat.completion", Created: created, Model: f.rt.p.Model,
		Usage: openai.Usage{PromptTokens: len(prompt), CompletionTokens: total, TotalTokens: len(prompt) + total}}
	for i := range states {
		c := openai.ChatCompletionChoice{Index: i, Message: openai.ChatCompletionMessage{Role: "assistant", Content: states[i].text}, FinishReason: states[i].reason}
		if req.LogProbs {
			c.LogProbs = &openai.LogProbs{Content: states[i].lps}
		}
		resp.Choices = append(resp.Choices, c)
	}
	w.Header().Set("Content-Type", "application/json")
	_ = json.NewEncoder(w).Encode(resp)
}

func (f *openAIFront) embeddings(w http.ResponseWriter, r *http.Request) {
	var req struct {
		Input json.RawMessage `json:"input"` // string | []string | [][]int (types/types.go:2707-2730)
		Model string          `json:"model"`
	}
	if err := json.NewDecoder(http.MaxBytesReader(w, r.Body, 10*1024*1024)).Decode(&req); err != nil {
		jsonError(w, http.StatusBadRequest, "invalid_request_error", err.Error())
		return
	}
	seqs, err := decodeEmbeddingInput(req.Input, f.tok)
	if err != nil {
		jsonError(w, http.StatusBadRequest, "invalid_request_error", err.Error())
		return
	}
	vecs, err := f.embed.Embed(seqs) // coalesced with the other requests in flight (embedBatcher)
	if err != nil {
		jsonError(w, http.StatusBadRequest, "invalid_request_error", err.Error())
		return
	}
	nTok := 0
	for _, s := range seqs {
		nTok += len(s)
	}
	resp := openai.EmbeddingResponse{Object: "list", Model: openai.EmbeddingModel(f.rt.p.Model), Usage: openai.Usage{PromptTokens: nTok, TotalTokens: nTok}}
	for i := range vecs {
		resp.Data = append(resp.Data, openai.Embedding{Object: "embedding", Index: i, Embedding: vecs[i]})
	}
	w.Header().Set("Content-Type", "application/json")
	_ = json.NewEncoder(w).Encode(resp)
}

// streamDecoder turns token ids into text for streaming.  Decoding each poll's ids alone garbles a character whose
// bytes / pieces straddle two polls, decoding the whole sequence at every poll is quadratic in the length of the answer:
// the ids not yet shown are decoded together with the previously shown batch as left context, that context's text is cut
// off the front, and an unfinished character (a trailing U+FFFD, at most its last three tokens) is held back unless the
// stream has ended.  Same algorithm as helix_b200/server.py StreamDecoder (fuzzed there against whole-sequence decoding).
type streamDecoder struct {
	tok     Tokenizer
	ids     []int32
	ctxOff  int // first id of the left context
	readOff int // ids before this index have been shown
}

func (d *streamDecoder) feed(ids []int32, final bool) string {
	d.ids = append(d.ids, ids...)
	n := len(d.ids)
	if d.readOff == n {
		return ""
	}
	ctx := d.tok.Decode(d.ids[d.ctxOff:d.readOff])
	cut := func(text string, upTo int) string {
		d.ctxOff, d.readOff = d.readOff, upTo
		if len(text) < len(ctx) {
			return ""
		}
		return text[len(ctx):]
	}
	maxHold := n - d.readOff
	if maxHold > 4 {
		maxHold = 4
	}
	if final {
		maxHold = 1
	}
	for hold := 0; hold < maxHold; hold++ {
		text := d.tok.Decode(d.ids[d.ctxOff : n-hold])
		if final || (!strings.HasSuffix(text, "\uFFFD") && len(text) > len(ctx)) {
			return cut(text, n-hold)
		}
	}
	if n-d.readOff > 8 { // not an unfinished character but invalid bytes: nothing later will repair them
		return cut(d.tok.Decode(d.ids[d.ctxOff:]), n)
	}
	return ""
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
