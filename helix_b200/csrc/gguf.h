// GGUF reader used by hb_model_load_gguf / hb_gguf_describe (gguf.cpp).
#pragma once
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/helix_b200.h"

namespace hb {

struct GgufTensor {
  std::string name;
  uint64_t ne[4];  // ne[0] is the contiguous dimension
  uint32_t type;   // ggml_type
  uint64_t offset; // relative to the data section
};

class GgufFile {
 public:
  ~GgufFile();
  bool open(const char* path, std::string* err);
  double num(const std::string& key, double dflt = 0) const;
  std::string str(const std::string& key) const;
  bool describe(hb_model_desc* d, std::string* why) const;
  bool read_f32(const std::string& name, std::vector<float>* out, size_t* rows, size_t* cols) const;
  const std::map<std::string, GgufTensor>& tensors() const { return tensors_; }

 private:
  const uint8_t* map_ = nullptr;
  size_t size_ = 0, data_off_ = 0;
  std::map<std::string, double> num_;
  std::map<std::string, std::string> str_;
  std::map<std::string, GgufTensor> tensors_;
};

std::string gguf_to_hf_name(const std::string& gguf_name);
void gguf_unpermute_rows(std::vector<float>& w, size_t rows, size_t cols, int n_head);
void gguf_to_bf16(const std::vector<float>& f, std::vector<uint16_t>* out);

}  // namespace hb
