// Package runner — nativeTokenizer: the Tokenizer (b200_front.go) backed by the library's hb_tok_* entry points
// (helix_b200/csrc/tokenizer.cpp: HF tokenizer.json, byte-level BPE, Llama-3 chat template; bit-exact with `tokenizers`
// 0.22 on that pipeline — tests/test_tokenizer_cpu.py).  helix_b200/tokenizer.py is the executable mirror of this file.
// Source only (no Go toolchain in this image).
package runner

/*
#include <stdlib.h>
#include "helix_b200.h"
*/
import "C"

import (
	"fmt"
	"unsafe"

	openai "github.com/sashabaranov/go-openai"
)

type nativeTokenizer struct {
	h   *C.hb_tokenizer
	eos int32
}

// NewNativeTokenizer loads <checkpoint dir>/tokenizer.json.
func NewNativeTokenizer(path string) (Tokenizer, error) {
	cpath := C.CString(path)
	defer C.free(unsafe.Pointer(cpath))
	t := &nativeTokenizer{}
	if rc := C.hb_tok_load(cpath, &t.h); rc != C.HB_OK {
		return nil, fmt.Errorf("helix-b200: hb_tok_load(%s): %d", path, int(rc))
	}
	eot := C.CString("<|eot_id|>")
	defer C.free(unsafe.Pointer(eot))
	t.eos = int32(C.hb_tok_token_id(t.h, eot))
	return t, nil
}

func (t *nativeTokenizer) Close() {
	if t.h != nil {
		C.hb_tok_free(t.h)
		t.h = nil
	}
}

func (t *nativeTokenizer) EOS() int32 { return t.eos }

// Encode: plain text, special-token strings inside it are NOT parsed (user content).
func (t *nativeTokenizer) Encode(text string) []int32 {
	ctext := C.CString(text)
	defer C.free(unsafe.Pointer(ctext))
	buf := make([]int32, len(text)+8) // a token covers at least one byte
	var n C.int32_t
	if rc := C.hb_tok_encode(t.h, ctext, 0, (*C.int32_t)(unsafe.Pointer(&buf[0])), C.int32_t(len(buf)), &n); rc != C.HB_OK {
		return nil
	}
	return buf[:int(n)]
}

func (t *nativeTokenizer) Decode(ids []int32) string {
	if len(ids) == 0 {
		return ""
	}
	out := make([]byte, 64*len(ids)+16)
	var ln C.size_t
	rc := C.hb_tok_decode(t.h, (*C.int32_t)(unsafe.Pointer(&ids[0])), C.int32_t(len(ids)), 1, (*C.char)(unsafe.Pointer(&out[0])), C.size_t(len(out)), &ln)
	if rc == C.HB_ERR_BUSY { // buffer too small: ln holds the needed length
		out = make([]byte, int(ln)+1)
		rc = C.hb_tok_decode(t.h, (*C.int32_t)(unsafe.Pointer(&ids[0])), C.int32_t(len(ids)), 1, (*C.char)(unsafe.Pointer(&out[0])), C.size_t(len(out)), &ln)
	}
	if rc != C.HB_OK {
		return ""
	}
	return string(out[:int(ln)]) // invalid UTF-8 tails (a character split across polls) are handled by emitStable
}

// EncodeChat renders the Llama-3 instruct template and tokenizes it in one piece (special tokens parsed), like
// HF's apply_chat_template followed by the tokenizer call.
func (t *nativeTokenizer) EncodeChat(messages []openai.ChatCompletionMessage) []int32 {
	n := len(messages)
	if n == 0 {
		return nil
	}
	roles := make([]*C.char, n)
	conts := make([]*C.char, n)
	total := 64
	for i, m := range messages {
		roles[i] = C.CString(m.Role)
		conts[i] = C.CString(m.Content)
		total += len(m.Content) + len(m.Role) + 16
	}
	defer func() {
		for i := range roles {
			C.free(unsafe.Pointer(roles[i]))
			C.free(unsafe.Pointer(conts[i]))
		}
	}()
	buf := make([]int32, total)
	var cnt C.int32_t
	rc := C.hb_tok_chat_llama3(t.h, (**C.char)(unsafe.Pointer(&roles[0])), (**C.char)(unsafe.Pointer(&conts[0])), C.int32_t(n),
		(*C.int32_t)(unsafe.Pointer(&buf[0])), C.int32_t(len(buf)), &cnt)
	if rc != C.HB_OK {
		return nil
	}
	return buf[:int(cnt)]
}
