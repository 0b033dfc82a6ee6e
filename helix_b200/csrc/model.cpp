#include "model.h"

#include <math.h>

namespace hb {

namespace {
constexpr size_t kAlign = 256;
size_t up(size_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

// Walks the tensors of a model in arena order. `visit(name, rows, cols, mode, group, kind)`;
// fused tensors are visited once per checkpoint tensor that lands in them.
struct Carver {
  uint8_t* base;
  size_t off = 0;
  bf16* take(size_t elems) {
    bf16* p = base ? reinterpret_cast<bf16*>(base + off) : nullptr;
    off += up(elems * sizeof(bf16));
    return p;
  }
};
}  // namespace

std::string validate_desc(const hb_model_desc& d) {
  if (d.arch != HB_ARCH_LLAMA && d.arch != HB_ARCH_BERT) return "unknown arch";
  if (d.hidden <= 0 || d.layers <= 0 || d.heads <= 0 || d.ffn <= 0 || d.vocab <= 0) return "non-positive dimension";
  if (d.head_dim != 64 && d.head_dim != 128) return "head_dim must be 64 or 128";
  if (d.hidden % 64) return "hidden must be a multiple of 64";
  if (d.hidden > 4096 * 4) return "hidden too large";
  if (d.arch == HB_ARCH_LLAMA) {
    if (d.kv_heads <= 0 || d.heads % d.kv_heads) return "heads must be a multiple of kv_heads";
    const int g = d.heads / d.kv_heads;
    if (g != 1 && g != 2 && g != 3 && g != 4 && g != 6 && g != 8) return "GQA group (heads/kv_heads) must be 1, 2, 3, 4, 6 or 8";
    if (d.ffn % 128) return "ffn must be a multiple of 128 (SwiGLU tile packing)";
    if (d.rope_theta <= 0.f) return "rope_theta must be positive";
    if (d.vocab % 4) return "vocab must be a multiple of 4 (16-byte rows of the fp32 logits)";
    if (d.hidden > 8192) return "hidden > 8192 unsupported (decode residual+RMSNorm register budget)";
  } else {
    if (d.heads * d.head_dim != d.hidden) return "BERT needs heads*head_dim == hidden";
    if (d.hidden > 4096) return "BERT hidden > 4096 unsupported (LayerNorm register budget)";
    if (d.ffn % 64) return "ffn must be a multiple of 64";
    if (d.max_pos <= 0) return "max_pos required";
  }
  return "";
}

static size_t carve(Model* m, const hb_model_desc& d, uint8_t* base) {
  Carver c{base};
  auto put = [&](const std::string& name, bf16* dst, size_t rows, size_t cols, int mode = 0, bool gain = false,
                 bool bias = false) {
    if (!m) return;
    Placement p;
    p.dst = dst;
    p.rows = rows;
    p.cols = cols;
    p.mode = mode;
    p.is_norm_gain = gain;
    p.is_bias = bias;
    m->placements[name] = p;
  };
  const size_t H = d.hidden, F = d.ffn, V = d.vocab, D = d.head_dim;
  if (d.arch == HB_ARCH_LLAMA) {
    const size_t QD = (size_t)d.heads * D, KD = (size_t)d.kv_heads * D;
    bf16* embed = c.take(V * H);
    put("model.embed_tokens.weight", embed, V, H);
    if (m) {
      m->embed = embed;
      m->ll.resize(d.layers);
    }
    for (int i = 0; i < d.layers; ++i) {
      const std::string p = "model.layers." + std::to_string(i) + ".";
      LlamaLayerW w;
      w.attn_norm = c.take(H);
      w.wqkv = c.take((QD + 2 * KD) * H);
      if (d.qkv_bias) w.bqkv = c.take(QD + 2 * KD);
      w.wo = c.take(H * QD);
      w.mlp_norm = c.take(H);
      w.wgu = c.take(2 * F * H);
      w.wdown = c.take(H * F);
      put(p + "input_layernorm.weight", w.attn_norm, H, 1, 0, true);
      put(p + "self_attn.q_proj.weight", w.wqkv, QD, H);
      put(p + "self_attn.k_proj.weight", w.wqkv ? w.wqkv + QD * H : nullptr, KD, H);
      put(p + "self_attn.v_proj.weight", w.wqkv ? w.wqkv + (QD + KD) * H : nullptr, KD, H);
      if (d.qkv_bias) {
        put(p + "self_attn.q_proj.bias", w.bqkv, QD, 1, 0, false, true);
        put(p + "self_attn.k_proj.bias", w.bqkv ? w.bqkv + QD : nullptr, KD, 1, 0, false, true);
        put(p + "self_attn.v_proj.bias", w.bqkv ? w.bqkv + QD + KD : nullptr, KD, 1, 0, false, true);
      }
      put(p + "self_attn.o_proj.weight", w.wo, H, QD);
      put(p + "post_attention_layernorm.weight", w.mlp_norm, H, 1, 0, true);
      put(p + "mlp.gate_proj.weight", w.wgu, F, H, 1);
      put(p + "mlp.up_proj.weight", w.wgu, F, H, 2);
      put(p + "mlp.down_proj.weight", w.wdown, H, F);
      if (m) m->ll[i] = w;
    }
    bf16* fn = c.take(H);
    put("model.norm.weight", fn, H, 1, 0, true);
    bf16* head = embed;
    if (!d.tie_embeddings) {
      head = c.take(V * H);
      put("lm_head.weight", head, V, H);
    }
    if (m) {
      m->final_norm = fn;
      m->lm_head = head;
    }
  } else {
    const size_t P = d.max_pos, TV = d.type_vocab > 0 ? d.type_vocab : 2;
    bf16* word = c.take(V * H);
    bf16* pos = c.take(P * H);
    bf16* type = c.take(TV * H);
    bf16* g = c.take(H);
    bf16* b = c.take(H);
    put("embeddings.word_embeddings.weight", word, V, H);
    put("embeddings.position_embeddings.weight", pos, P, H);
    put("embeddings.token_type_embeddings.weight", type, TV, H);
    put("embeddings.LayerNorm.weight", g, H, 1, 0, true);
    put("embeddings.LayerNorm.bias", b, H, 1, 0, false, true);
    if (m) {
      m->word = word;
      m->pos = pos;
      m->type = type;
      m->emb_ln_g = g;
      m->emb_ln_b = b;
      m->bl.resize(d.layers);
    }
    for (int i = 0; i < d.layers; ++i) {
      const std::string p = "encoder.layer." + std::to_string(i) + ".";
      BertLayerW w;
      w.wqkv = c.take(3 * H * H);
      w.bqkv = c.take(3 * H);
      w.wo = c.take(H * H);
      w.bo = c.take(H);
      w.ln1_g = c.take(H);
      w.ln1_b = c.take(H);
      w.w1 = c.take(F * H);
      w.b1 = c.take(F);
      w.w2 = c.take(H * F);
      w.b2 = c.take(H);
      w.ln2_g = c.take(H);
      w.ln2_b = c.take(H);
      put(p + "attention.self.query.weight", w.wqkv, H, H);
      put(p + "attention.self.key.weight", w.wqkv ? w.wqkv + H * H : nullptr, H, H);
      put(p + "attention.self.value.weight", w.wqkv ? w.wqkv + 2 * H * H : nullptr, H, H);
      put(p + "attention.self.query.bias", w.bqkv, H, 1, 0, false, true);
      put(p + "attention.self.key.bias", w.bqkv ? w.bqkv + H : nullptr, H, 1, 0, false, true);
      put(p + "attention.self.value.bias", w.bqkv ? w.bqkv + 2 * H : nullptr, H, 1, 0, false, true);
      put(p + "attention.output.dense.weight", w.wo, H, H);
      put(p + "attention.output.dense.bias", w.bo, H, 1, 0, false, true);
      put(p + "attention.output.LayerNorm.weight", w.ln1_g, H, 1, 0, true);
      put(p + "attention.output.LayerNorm.bias", w.ln1_b, H, 1, 0, false, true);
      put(p + "intermediate.dense.weight", w.w1, F, H);
      put(p + "intermediate.dense.bias", w.b1, F, 1, 0, false, true);
      put(p + "output.dense.weight", w.w2, H, F);
      put(p + "output.dense.bias", w.b2, H, 1, 0, false, true);
      put(p + "output.LayerNorm.weight", w.ln2_g, H, 1, 0, true);
      put(p + "output.LayerNorm.bias", w.ln2_b, H, 1, 0, false, true);
      if (m) m->bl[i] = w;
    }
  }
  return c.off;
}

size_t arena_bytes_for(const hb_model_desc& d) { return carve(nullptr, d, nullptr); }

void layout_model(Model& m) {
  m.placements.clear();
  m.filled.clear();
  carve(&m, m.d, reinterpret_cast<uint8_t*>(m.arena));
}

std::vector<float> rope_inv_freq(const hb_model_desc& d) {
  const int dim = d.head_dim;
  std::vector<float> f(dim / 2);
  for (int i = 0; i < dim / 2; ++i) {
    // HF: 1.0 / (base ** (arange(0, dim, 2).float() / dim)), evaluated in fp32
    const float e = (float)(2 * i) / (float)dim;
    f[i] = 1.0f / powf(d.rope_theta, e);
  }
  if (d.rope_factor > 0.f) {
    const float factor = d.rope_factor, lo = d.rope_low_freq_factor, hi = d.rope_high_freq_factor;
    const float old_len = (float)d.rope_orig_max_pos;
    const float low_wl = old_len / lo, high_wl = old_len / hi;
    for (int i = 0; i < dim / 2; ++i) {
      const float inv = f[i];
      const float wl = 2.0f * (float)M_PI / inv;
      float v = (wl > low_wl) ? inv / factor : inv;
      const float smooth = (old_len / wl - lo) / (hi - lo);
      const float smoothed = (1.0f - smooth) * v / factor + smooth * v;
      const bool medium = !(wl < high_wl) && !(wl > low_wl);
      f[i] = medium ? smoothed : v;
    }
  }
  return f;
}

}  // namespace hb
