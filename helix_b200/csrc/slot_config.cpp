// hb_slot_config: the control-plane side of the wire protocol in compiled code — decodes the JSON the scheduler sends
// to the runner to create a slot (types.CreateRunnerSlotRequest / CreateRunnerSlotAttributes, api/pkg/types/runner.go:
// 92-109) and derives the engine configuration the way Slot.Create + VLLMRuntime do for a vLLM-style slot
// (api/pkg/runner/slot.go:395-470: model override and the three accepted shapes of runtime_args.args;
// api/pkg/runner/vllm_runtime.go:705-762: the flags that reach the backend; api/pkg/scheduler/runner.go:1187-1259: how the
// scheduler computes --gpu-memory-utilization).  helix_b200/runtime.py is the Python mirror; tests/test_slot_config_cpu.py
// checks the two against each other and against the reference's ratio known-answer cases.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "../../include/helix_b200.h"
#include "json_min.h"

namespace {
using hbjson::JVal;

// fmt.Sprintf("%v", v) of a JSON scalar as Go prints it after encoding/json decoded it into interface{} (numbers are
// float64: integral values print without a fraction)
std::string go_print(const JVal& v) {
  if (v.kind == JVal::NUM) {
    if (v.num == floor(v.num) && fabs(v.num) < 1e15) {
      char b[32];
      snprintf(b, sizeof b, "%lld", (long long)v.num);
      return b;
    }
    char b[32];
    snprintf(b, sizeof b, "%g", v.num);
    return b;
  }
  return v.raw;
}

}  // namespace

extern "C" int hb_slot_config(const char* json, uint64_t per_gpu_memory_bytes, hb_engine_cfg* cfg, hb_slot_info* info) {
  if (!json || !cfg || !info) return HB_ERR_INVALID;
  JVal root;
  if (!hbjson::parse(json, strlen(json), &root) || root.kind != JVal::OBJ) return HB_ERR_INVALID;
  const JVal* at = root.get("attributes");  // a CreateRunnerSlotRequest wraps the attributes; accept both
  if (!at) at = &root;
  if (at->kind != JVal::OBJ) return HB_ERR_INVALID;
  memset(cfg, 0, sizeof *cfg);
  memset(info, 0, sizeof *info);
  const JVal* rt = at->get("runtime");
  if (!rt || rt->kind != JVal::STR || rt->raw != "vllm") return HB_ERR_INVALID;  // plug-in option A: vLLM-style slots only
  std::string model;
  if (const JVal* m = at->get("model"); m && m->kind == JVal::STR) model = m->raw;
  uint64_t mem_req = 0;
  if (const JVal* m = at->get("model_memory_requirement"); m && m->kind == JVal::NUM && m->num > 0) mem_req = (uint64_t)m->num;
  long long ctx_len = 0;
  if (const JVal* c = at->get("context_length"); c && c->kind == JVal::NUM) ctx_len = (long long)c->num;
  if (const JVal* g = at->get("gpu_index"); g && g->kind == JVal::NUM) cfg->device = (int32_t)g->num;
  if (const JVal* t = at->get("tensor_parallel_size"); t && t->kind == JVal::NUM) info->tensor_parallel_size = (int32_t)t->num;

  std::vector<std::string> args;
  if (const JVal* ra = at->get("runtime_args"); ra && ra->kind == JVal::OBJ) {
    if (const JVal* m = ra->get("model"); m && m->kind == JVal::STR && !m->raw.empty()) model = m->raw;  // slot.go:410
    if (const JVal* a = ra->get("args")) {
      if (a->kind == JVal::ARR) {                    // []string or []interface{} (slot.go:418-434)
        for (const JVal& v : a->arr) args.push_back(go_print(v));
      } else if (a->kind == JVal::OBJ) {             // map form -> ["--k", "v", ...] (slot.go:435-447)
        for (const auto& kv : a->obj) {
          args.push_back(kv.first.rfind("--", 0) == 0 ? kv.first : "--" + kv.first);
          args.push_back(go_print(kv.second));
        }
      }
    }
  }
  if (model.empty()) return HB_ERR_INVALID;  // "model must be specified for vLLM runtime" (slot.go:463-466)
  snprintf(info->model, sizeof info->model, "%s", model.c_str());

  // the flags the scheduler emits for a slot (helix_b200/runtime.py parse_vllm_args is the mirror)
  int max_seqs = 256, max_model_len = 0, max_batched = 0, prefix = 1;
  double util = 0;
  for (size_t i = 0; i < args.size();) {
    const std::string& a = args[i];
    const bool has = i + 1 < args.size();
    if (a == "--gpu-memory-utilization" && has) { util = atof(args[i + 1].c_str()); i += 2; }
    else if (a == "--max-num-seqs" && has) { max_seqs = atoi(args[i + 1].c_str()); i += 2; }
    else if (a == "--max-model-len" && has) { max_model_len = atoi(args[i + 1].c_str()); i += 2; }
    else if (a == "--task" && has) { info->is_embed = args[i + 1] == "embed"; i += 2; }
    else if (a == "--max-num-batched-tokens" && has) { max_batched = atoi(args[i + 1].c_str()); i += 2; }
    else if (a == "--enable-prefix-caching") { prefix = 1; i += 1; }
    else if (a == "--no-enable-prefix-caching") { prefix = 0; i += 1; }
    else if (a.rfind("--", 0) == 0 && has && args[i + 1].rfind("--", 0) != 0) { info->n_unknown_args += 2; i += 2; }
    else { info->n_unknown_args += 1; i += 1; }
  }
  info->gpu_memory_utilization = (float)util;
  // budget: the exact bytes the scheduler packed the slot with; the ratio flag is the fallback
  cfg->memory_budget_bytes = mem_req ? mem_req : (util > 0 && per_gpu_memory_bytes ? (uint64_t)((double)per_gpu_memory_bytes * util) : 0);
  cfg->max_seqs = max_seqs;
  cfg->max_ctx = max_model_len > 0 ? max_model_len : (ctx_len > 0 ? (int32_t)ctx_len : 0);
  cfg->max_batched_tokens = max_batched;
  cfg->kv_page_size = 64;
  cfg->use_cuda_graphs = 1;
  cfg->enable_prefix_cache = (prefix && !info->is_embed) ? 1 : 0;
  cfg->decode_with_prefill = 1;
  return HB_OK;
}
