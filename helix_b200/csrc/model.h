// Weight arena: every tensor of a ModelInstance lives in ONE device allocation (so a replica can be
// filled by a single NCCL broadcast — SURVEY.md §8e) with projections pre-fused the way the kernels
// consume them:  wqkv = [q;k;v] rows,  w_gate_up packed per 256-row tile as [128 gate | 128 up].
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/helix_b200.h"
#include "kernels.h"

namespace hb {

struct LlamaLayerW {
  bf16 *attn_norm, *wqkv, *wo, *mlp_norm, *wgu, *wdown;
  bf16* bqkv = nullptr;  // [q;k;v] projection biases (hb_model_desc.qkv_bias), else null
};
struct BertLayerW {
  bf16 *wqkv, *bqkv, *wo, *bo, *ln1_g, *ln1_b, *w1, *b1, *w2, *b2, *ln2_g, *ln2_b;
};

// where an uploaded checkpoint tensor lands inside the arena
struct Placement {
  bf16* dst = nullptr;
  size_t rows = 0, cols = 0;   // checkpoint tensor shape (cols == 1 for vectors)
  int mode = 0;                // 0: contiguous copy; 1: gate rows -> tile*256 + r%128; 2: up rows -> tile*256+128 + r%128
  bool is_norm_gain = false;   // random init sets these to 1
  bool is_bias = false;        // random init keeps these small
};

struct Model {
  hb_model_desc d{};
  bf16* arena = nullptr;
  size_t arena_bytes = 0;
  float* inv_freq = nullptr;  // [head_dim/2] device (llama)
  // llama
  bf16 *embed = nullptr, *final_norm = nullptr, *lm_head = nullptr;
  std::vector<LlamaLayerW> ll;
  // bert
  bf16 *word = nullptr, *pos = nullptr, *type = nullptr, *emb_ln_g = nullptr, *emb_ln_b = nullptr;
  std::vector<BertLayerW> bl;
  std::unordered_map<std::string, Placement> placements;  // HF checkpoint name -> arena slot
  std::unordered_map<std::string, bool> filled;

  int qkv_cols() const { return (d.heads + 2 * d.kv_heads) * d.head_dim; }
};

// Validates the description against what the kernels support; returns "" when fine.
std::string validate_desc(const hb_model_desc& d);
// Arena size in bytes for a description (also used by hb_memory_estimate).
size_t arena_bytes_for(const hb_model_desc& d);
// Carves `m.arena` (already allocated, m.arena_bytes long) into tensors and fills m.placements.
void layout_model(Model& m);
// Host inverse frequencies incl. llama3 scaling (HF transformers `_compute_llama3_parameters`).
std::vector<float> rope_inv_freq(const hb_model_desc& d);

}  // namespace hb
