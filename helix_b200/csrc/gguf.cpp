// GGUF checkpoints (scope row F3): the reference's Ollama backend serves llama.cpp GGUF blobs — its catalogue entry
// "llama3:instruct" is an 8B Q4_0 file (api/pkg/model/models.go:259-266; blobs under the Ollama cache,
// api/pkg/runner/runner_cmd.go:213-249) and its memory estimator parses the same files (api/pkg/memory/gguf.go:35-170).
// This reader takes such a file straight into the engine: header + metadata -> hb_model_desc, every tensor dequantised to
// bf16 (F32 / F16 / BF16 / Q8_0 / Q4_0 / Q4_1 / Q5_0 / Q5_1 / Q4_K / Q5_K / Q6_K — the block formats as published in ggml's
// ggml-quants.c, restated here and in tests/gguf_ref.py), tensor names mapped to the HF checkpoint names the engine's arena
// uses, and the row permutation llama.cpp's converter applies to q/k projections (interleaved-pair RoPE) undone, because
// the kernels use the HF rotate-half convention.
#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/helix_b200.h"
#include "gguf.h"

namespace hb {
namespace {

float f16_to_f32(uint16_t h) {
  const uint32_t sign = (uint32_t)(h >> 15) << 31, exp = (h >> 10) & 0x1F, man = h & 0x3FF;
  uint32_t u;
  if (exp == 0) {
    if (man == 0) { u = sign; }
    else {
      uint32_t e = 127 - 15 + 1, m = man;
      while (!(m & 0x400)) { m <<= 1; --e; }
      u = sign | (e << 23) | ((m & 0x3FF) << 13);
    }
  } else if (exp == 0x1F) { u = sign | 0x7F800000u | (man << 13); }
  else { u = sign | ((exp + 127 - 15) << 23) | (man << 13); }
  float f;
  memcpy(&f, &u, 4);
  return f;
}
uint16_t f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u && (u & 0x7FFFFFu)) return (uint16_t)((u >> 16) | 0x40);  // NaN stays NaN
  u += ((u >> 16) & 1u) + 0x7FFFu;
  return (uint16_t)(u >> 16);
}
uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }

struct TypeInfo { int block; int bytes; };
bool type_info(uint32_t t, TypeInfo* o) {
  switch (t) {
    case 0: *o = {1, 4}; return true;      // F32
    case 1: *o = {1, 2}; return true;      // F16
    case 2: *o = {32, 18}; return true;    // Q4_0
    case 3: *o = {32, 20}; return true;    // Q4_1
    case 6: *o = {32, 22}; return true;    // Q5_0
    case 7: *o = {32, 24}; return true;    // Q5_1
    case 8: *o = {32, 34}; return true;    // Q8_0
    case 12: *o = {256, 144}; return true; // Q4_K
    case 13: *o = {256, 176}; return true; // Q5_K
    case 14: *o = {256, 210}; return true; // Q6_K
    case 30: *o = {1, 2}; return true;     // BF16
    default: return false;
  }
}

void scale_min_k4(int j, const uint8_t* q, uint8_t* d, uint8_t* m) {
  if (j < 4) { *d = q[j] & 63; *m = q[j + 4] & 63; }
  else { *d = (q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4); *m = (q[j + 4] >> 4) | ((q[j] >> 6) << 4); }
}

// one block of `type` at `b` -> y[block elements]  (ggml-quants.c dequantize_row_*)
void dequant_block(uint32_t type, const uint8_t* b, float* y) {
  switch (type) {
    case 2: {  // Q4_0: fp16 d, 16 bytes of nibbles
      const float d = f16_to_f32(rd16(b));
      const uint8_t* qs = b + 2;
      for (int j = 0; j < 16; ++j) { y[j] = ((qs[j] & 0xF) - 8) * d; y[j + 16] = ((qs[j] >> 4) - 8) * d; }
      break;
    }
    case 3: {  // Q4_1: fp16 d, fp16 m
      const float d = f16_to_f32(rd16(b)), m = f16_to_f32(rd16(b + 2));
      const uint8_t* qs = b + 4;
      for (int j = 0; j < 16; ++j) { y[j] = (qs[j] & 0xF) * d + m; y[j + 16] = (qs[j] >> 4) * d + m; }
      break;
    }
    case 6: {  // Q5_0: fp16 d, 32 high bits, 16 bytes of nibbles
      const float d = f16_to_f32(rd16(b));
      uint32_t qh;
      memcpy(&qh, b + 2, 4);
      const uint8_t* qs = b + 6;
      for (int j = 0; j < 16; ++j) {
        const uint8_t h0 = ((qh >> j) << 4) & 0x10, h1 = (qh >> (j + 12)) & 0x10;
        y[j] = (int)(((qs[j] & 0xF) | h0) - 16) * d;
        y[j + 16] = (int)(((qs[j] >> 4) | h1) - 16) * d;
      }
      break;
    }
    case 7: {  // Q5_1: fp16 d, fp16 m, 32 high bits, nibbles
      const float d = f16_to_f32(rd16(b)), m = f16_to_f32(rd16(b + 2));
      uint32_t qh;
      memcpy(&qh, b + 4, 4);
      const uint8_t* qs = b + 8;
      for (int j = 0; j < 16; ++j) {
        const uint8_t h0 = ((qh >> j) << 4) & 0x10, h1 = (qh >> (j + 12)) & 0x10;
        y[j] = ((qs[j] & 0xF) | h0) * d + m;
        y[j + 16] = ((qs[j] >> 4) | h1) * d + m;
      }
      break;
    }
    case 8: {  // Q8_0: fp16 d, 32 int8
      const float d = f16_to_f32(rd16(b));
      const int8_t* qs = reinterpret_cast<const int8_t*>(b + 2);
      for (int j = 0; j < 32; ++j) y[j] = qs[j] * d;
      break;
    }
    case 12: {  // Q4_K: fp16 d, fp16 dmin, 12 bytes of 6-bit scales/mins, 128 bytes of nibbles
      const float d = f16_to_f32(rd16(b)), dmin = f16_to_f32(rd16(b + 2));
      const uint8_t* sc = b + 4;
      const uint8_t* q = b + 16;
      int is = 0;
      for (int j = 0; j < 256; j += 64) {
        uint8_t s, m;
        scale_min_k4(is + 0, sc, &s, &m);
        const float d1 = d * s, m1 = dmin * m;
        scale_min_k4(is + 1, sc, &s, &m);
        const float d2 = d * s, m2 = dmin * m;
        for (int l = 0; l < 32; ++l) y[j + l] = d1 * (q[l] & 0xF) - m1;
        for (int l = 0; l < 32; ++l) y[j + 32 + l] = d2 * (q[l] >> 4) - m2;
        q += 32;
        is += 2;
      }
      break;
    }
    case 13: {  // Q5_K: d, dmin, scales[12], qh[32], qs[128]
      const float d = f16_to_f32(rd16(b)), dmin = f16_to_f32(rd16(b + 2));
      const uint8_t* sc = b + 4;
      const uint8_t* qh = b + 16;
      const uint8_t* ql = b + 48;
      int is = 0;
      uint8_t u1 = 1, u2 = 2;
      for (int j = 0; j < 256; j += 64) {
        uint8_t s, m;
        scale_min_k4(is + 0, sc, &s, &m);
        const float d1 = d * s, m1 = dmin * m;
        scale_min_k4(is + 1, sc, &s, &m);
        const float d2 = d * s, m2 = dmin * m;
        for (int l = 0; l < 32; ++l) y[j + l] = d1 * ((ql[l] & 0xF) + (qh[l] & u1 ? 16 : 0)) - m1;
        for (int l = 0; l < 32; ++l) y[j + 32 + l] = d2 * ((ql[l] >> 4) + (qh[l] & u2 ? 16 : 0)) - m2;
        ql += 32;
        is += 2;
        u1 <<= 2;
        u2 <<= 2;
      }
      break;
    }
    case 14: {  // Q6_K: ql[128], qh[64], int8 scales[16], fp16 d
      const uint8_t* ql = b;
      const uint8_t* qh = b + 128;
      const int8_t* sc = reinterpret_cast<const int8_t*>(b + 192);
      const float d = f16_to_f32(rd16(b + 208));
      for (int n = 0; n < 256; n += 128) {
        for (int l = 0; l < 32; ++l) {
          const int is = l / 16;
          const int q1 = (int)((ql[l] & 0xF) | (((qh[l] >> 0) & 3) << 4)) - 32;
          const int q2 = (int)((ql[l + 32] & 0xF) | (((qh[l] >> 2) & 3) << 4)) - 32;
          const int q3 = (int)((ql[l] >> 4) | (((qh[l] >> 4) & 3) << 4)) - 32;
          const int q4 = (int)((ql[l + 32] >> 4) | (((qh[l] >> 6) & 3) << 4)) - 32;
          y[n + l] = d * sc[is + 0] * q1;
          y[n + l + 32] = d * sc[is + 2] * q2;
          y[n + l + 64] = d * sc[is + 4] * q3;
          y[n + l + 96] = d * sc[is + 6] * q4;
        }
        ql += 64;
        qh += 32;
        sc += 8;
      }
      break;
    }
    default: break;
  }
}

struct Cursor {
  const uint8_t* p;
  const uint8_t* end;
  bool ok = true;
  template <typename T>
  T get() {
    T v{};
    if (p + sizeof(T) > end) { ok = false; return v; }
    memcpy(&v, p, sizeof(T));
    p += sizeof(T);
    return v;
  }
  std::string str() {
    const uint64_t n = get<uint64_t>();
    if (!ok || n > (uint64_t)(end - p)) { ok = false; return ""; }
    std::string s(reinterpret_cast<const char*>(p), (size_t)n);
    p += n;
    return s;
  }
};

}  // namespace

GgufFile::~GgufFile() {
  if (map_) munmap(const_cast<uint8_t*>(map_), size_);
}

bool GgufFile::open(const char* path, std::string* err) {
  const int fd = ::open(path, O_RDONLY);
  if (fd < 0) { *err = std::string("cannot open ") + path; return false; }
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size < 24) { close(fd); *err = "not a GGUF file (too small)"; return false; }
  size_ = (size_t)st.st_size;
  void* m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (m == MAP_FAILED) { *err = "mmap failed"; return false; }
  map_ = static_cast<const uint8_t*>(m);
  Cursor c{map_, map_ + size_};
  if (c.get<uint32_t>() != 0x46554747u) { *err = "bad magic (not GGUF)"; return false; }
  const uint32_t version = c.get<uint32_t>();
  if (version < 2 || version > 3) { *err = "unsupported GGUF version " + std::to_string(version); return false; }
  const uint64_t n_tensors = c.get<uint64_t>(), n_kv = c.get<uint64_t>();
  if (n_tensors > (1u << 20) || n_kv > (1u << 20)) { *err = "implausible GGUF header"; return false; }
  auto scalar = [&](uint32_t type, double* num, std::string* s) -> bool {
    switch (type) {
      case 0: *num = c.get<uint8_t>(); return true;
      case 1: *num = c.get<int8_t>(); return true;
      case 2: *num = c.get<uint16_t>(); return true;
      case 3: *num = c.get<int16_t>(); return true;
      case 4: *num = c.get<uint32_t>(); return true;
      case 5: *num = c.get<int32_t>(); return true;
      case 6: *num = c.get<float>(); return true;
      case 7: *num = c.get<uint8_t>(); return true;
      case 8: *s = c.str(); return true;
      case 10: *num = (double)c.get<uint64_t>(); return true;
      case 11: *num = (double)c.get<int64_t>(); return true;
      case 12: *num = c.get<double>(); return true;
      default: return false;
    }
  };
  for (uint64_t i = 0; i < n_kv && c.ok; ++i) {
    const std::string key = c.str();
    const uint32_t type = c.get<uint32_t>();
    if (type == 9) {  // array: only its length is kept (vocabulary tables are not needed here)
      const uint32_t et = c.get<uint32_t>();
      const uint64_t cnt = c.get<uint64_t>();
      for (uint64_t k = 0; k < cnt && c.ok; ++k) {
        double num;
        std::string s;
        if (!scalar(et, &num, &s)) { *err = "unsupported array element type in " + key; return false; }
      }
      num_[key + ".length"] = (double)cnt;
    } else {
      double num = 0;
      std::string s;
      if (!scalar(type, &num, &s)) { *err = "unsupported metadata type for " + key; return false; }
      if (type == 8) str_[key] = s; else num_[key] = num;
    }
  }
  for (uint64_t i = 0; i < n_tensors && c.ok; ++i) {
    GgufTensor t;
    t.name = c.str();
    const uint32_t nd = c.get<uint32_t>();
    if (nd > 4) { *err = "tensor with more than 4 dimensions"; return false; }
    t.ne[0] = t.ne[1] = t.ne[2] = t.ne[3] = 1;
    for (uint32_t k = 0; k < nd; ++k) t.ne[k] = c.get<uint64_t>();
    t.type = c.get<uint32_t>();
    t.offset = c.get<uint64_t>();
    tensors_[t.name] = t;
  }
  if (!c.ok) { *err = "truncated GGUF header"; return false; }
  size_t align = 32;
  if (auto it = num_.find("general.alignment"); it != num_.end() && it->second >= 1) align = (size_t)it->second;
  data_off_ = ((size_t)(c.p - map_) + align - 1) / align * align;
  for (auto& kv : tensors_) {
    const GgufTensor& t = kv.second;
    TypeInfo ti;
    if (!type_info(t.type, &ti)) { *err = "tensor " + t.name + ": ggml type " + std::to_string(t.type) + " not supported"; return false; }
    if (t.ne[0] % ti.block) { *err = "tensor " + t.name + ": row length not a multiple of the block size"; return false; }
    const size_t bytes = (size_t)(t.ne[0] / ti.block) * ti.bytes * t.ne[1] * t.ne[2] * t.ne[3];
    if (data_off_ + t.offset + bytes > size_) { *err = "tensor " + t.name + " runs past the end of the file"; return false; }
  }
  return true;
}

double GgufFile::num(const std::string& key, double dflt) const {
  auto it = num_.find(key);
  return it == num_.end() ? dflt : it->second;
}
std::string GgufFile::str(const std::string& key) const {
  auto it = str_.find(key);
  return it == str_.end() ? "" : it->second;
}

// rows x cols fp32, row-major, of a 1-D / 2-D tensor (ne[0] = cols is the contiguous dimension)
bool GgufFile::read_f32(const std::string& name, std::vector<float>* out, size_t* rows, size_t* cols) const {
  auto it = tensors_.find(name);
  if (it == tensors_.end()) return false;
  const GgufTensor& t = it->second;
  TypeInfo ti;
  type_info(t.type, &ti);
  *cols = (size_t)t.ne[0];
  *rows = (size_t)(t.ne[1] * t.ne[2] * t.ne[3]);
  out->resize(*rows * *cols);
  const uint8_t* src = map_ + data_off_ + t.offset;
  const size_t n = out->size();
  float* y = out->data();
  if (t.type == 0) { memcpy(y, src, n * 4); return true; }
  if (t.type == 1) { for (size_t i = 0; i < n; ++i) y[i] = f16_to_f32(rd16(src + 2 * i)); return true; }
  if (t.type == 30) {
    for (size_t i = 0; i < n; ++i) { const uint32_t u = (uint32_t)rd16(src + 2 * i) << 16; memcpy(&y[i], &u, 4); }
    return true;
  }
  const size_t blocks = n / ti.block;
  for (size_t b = 0; b < blocks; ++b) dequant_block(t.type, src + b * ti.bytes, y + b * ti.block);
  return true;
}

// metadata -> model description; *why says what is unsupported
bool GgufFile::describe(hb_model_desc* d, std::string* why) const {
  memset(d, 0, sizeof *d);
  const std::string arch = str("general.architecture");
  if (arch != "llama" && arch != "qwen2") { *why = "general.architecture '" + arch + "' is not served (llama, qwen2)"; return false; }
  const std::string p = arch + ".";
  d->arch = HB_ARCH_LLAMA;
  d->hidden = (int32_t)num(p + "embedding_length");
  d->layers = (int32_t)num(p + "block_count");
  d->heads = (int32_t)num(p + "attention.head_count");
  d->kv_heads = (int32_t)num(p + "attention.head_count_kv", d->heads);
  d->ffn = (int32_t)num(p + "feed_forward_length");
  d->head_dim = (int32_t)num(p + "attention.key_length", d->heads ? d->hidden / d->heads : 0);
  d->max_pos = (int32_t)num(p + "context_length", 8192);
  d->norm_eps = (float)num(p + "attention.layer_norm_rms_epsilon", 1e-5);
  d->rope_theta = (float)num(p + "rope.freq_base", 10000.0);
  d->rope_low_freq_factor = 1.f;
  d->rope_high_freq_factor = 4.f;
  d->qkv_bias = arch == "qwen2" ? 1 : 0;
  auto emb = tensors_.find("token_embd.weight");
  if (emb == tensors_.end()) { *why = "token_embd.weight missing"; return false; }
  d->vocab = (int32_t)emb->second.ne[1];
  d->tie_embeddings = tensors_.count("output.weight") ? 0 : 1;
  if (tensors_.count("rope_freqs.weight")) { *why = "rope_freqs.weight (llama3 frequency factors) is not supported: use the HF checkpoint with its rope_scaling"; return false; }
  if (num(p + "rope.scaling.factor", 0) > 0 && str(p + "rope.scaling.type") != "none" && !str(p + "rope.scaling.type").empty()) {
    *why = "rope scaling type '" + str(p + "rope.scaling.type") + "' is not supported";
    return false;
  }
  return true;
}

// HF checkpoint name of a GGUF tensor (llama / qwen2 families); "" = not used by the engine
std::string gguf_to_hf_name(const std::string& g) {
  if (g == "token_embd.weight") return "model.embed_tokens.weight";
  if (g == "output_norm.weight") return "model.norm.weight";
  if (g == "output.weight") return "lm_head.weight";
  if (g.rfind("blk.", 0) != 0) return "";
  const size_t dot = g.find('.', 4);
  if (dot == std::string::npos) return "";
  const std::string layer = g.substr(4, dot - 4), rest = g.substr(dot + 1);
  static const std::map<std::string, std::string> m = {
      {"attn_norm.weight", "input_layernorm.weight"},       {"attn_q.weight", "self_attn.q_proj.weight"},
      {"attn_k.weight", "self_attn.k_proj.weight"},         {"attn_v.weight", "self_attn.v_proj.weight"},
      {"attn_q.bias", "self_attn.q_proj.bias"},             {"attn_k.bias", "self_attn.k_proj.bias"},
      {"attn_v.bias", "self_attn.v_proj.bias"},             {"attn_output.weight", "self_attn.o_proj.weight"},
      {"ffn_norm.weight", "post_attention_layernorm.weight"}, {"ffn_gate.weight", "mlp.gate_proj.weight"},
      {"ffn_up.weight", "mlp.up_proj.weight"},              {"ffn_down.weight", "mlp.down_proj.weight"}};
  auto it = m.find(rest);
  return it == m.end() ? "" : "model.layers." + layer + "." + it->second;
}

// llama.cpp's converter stores q/k projection rows permuted for its interleaved-pair RoPE:
//   gguf_row[h][2*i + p] = hf_row[h][p * D/2 + i]   (p = 0,1: the two halves rotate_half pairs up).  Undo it.
void gguf_unpermute_rows(std::vector<float>& w, size_t rows, size_t cols, int n_head) {
  if (n_head <= 0 || rows % (size_t)n_head) return;
  const size_t D = rows / (size_t)n_head, half = D / 2;
  std::vector<float> out(w.size());
  for (int h = 0; h < n_head; ++h)
    for (size_t i = 0; i < half; ++i)
      for (int p = 0; p < 2; ++p)
        memcpy(&out[((size_t)h * D + (size_t)p * half + i) * cols], &w[((size_t)h * D + 2 * i + (size_t)p) * cols], cols * sizeof(float));
  w.swap(out);
}

void gguf_to_bf16(const std::vector<float>& f, std::vector<uint16_t>* out) {
  out->resize(f.size());
  for (size_t i = 0; i < f.size(); ++i) (*out)[i] = f32_to_bf16(f[i]);
}

}  // namespace hb
