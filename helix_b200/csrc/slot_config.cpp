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

namespace {

struct JVal {
  enum Kind { NUL, BOOL, NUM, STR, ARR, OBJ } kind = NUL;
  bool b = false;
  double num = 0;
  std::string raw;  // string value, or the literal text of a number (so 256 prints as "256", 0.5 as "0.5")
  std::vector<JVal> arr;
  std::vector<std::pair<std::string, JVal>> obj;  // document order
  const JVal* get(const char* k) const {
    if (kind != OBJ) return nullptr;
    for (const auto& kv : obj)
      if (kv.first == k) return &kv.second;
    return nullptr;
  }
};

struct Parser {
  const char* p;
  const char* end;
  bool ok = true;
  void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
  bool lit(const char* s) {
    const size_t n = strlen(s);
    if ((size_t)(end - p) >= n && !strncmp(p, s, n)) { p += n; return true; }
    return false;
  }
  static void utf8(std::string& o, unsigned cp) {
    if (cp < 0x80) o += (char)cp;
    else if (cp < 0x800) { o += (char)(0xC0 | (cp >> 6)); o += (char)(0x80 | (cp & 0x3F)); }
    else if (cp < 0x10000) { o += (char)(0xE0 | (cp >> 12)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
    else { o += (char)(0xF0 | (cp >> 18)); o += (char)(0x80 | ((cp >> 12) & 0x3F)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
  }
  bool str(std::string& out) {
    if (p >= end || *p != '"') return ok = false;
    ++p;
    while (p < end && *p != '"') {
      if (*p == '\\') {
        if (++p >= end) return ok = false;
        switch (*p) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'u': {
            if (end - p < 5) return ok = false;
            unsigned cp = (unsigned)strtoul(std::string(p + 1, 4).c_str(), nullptr, 16);
            p += 4;
            if (cp >= 0xD800 && cp < 0xDC00 && end - p >= 7 && p[1] == '\\' && p[2] == 'u') {  // surrogate pair
              const unsigned lo = (unsigned)strtoul(std::string(p + 3, 4).c_str(), nullptr, 16);
              cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
              p += 6;
            }
            utf8(out, cp);
            break;
          }
          default: out += *p;  // \" \\ \/
        }
        ++p;
      } else {
        out += *p++;
      }
    }
    if (p >= end) return ok = false;
    ++p;
    return true;
  }
  JVal val(int depth = 0) {
    JVal v;
    ws();
    if (!ok || p >= end || depth > 64) { ok = false; return v; }
    if (*p == '{') {
      v.kind = JVal::OBJ;
      ++p;
      ws();
      if (p < end && *p == '}') { ++p; return v; }
      while (ok) {
        ws();
        std::string k;
        if (!str(k)) break;
        ws();
        if (p >= end || *p != ':') { ok = false; break; }
        ++p;
        v.obj.emplace_back(std::move(k), val(depth + 1));
        ws();
        if (p < end && *p == ',') { ++p; continue; }
        if (p < end && *p == '}') { ++p; break; }
        ok = false;
      }
    } else if (*p == '[') {
      v.kind = JVal::ARR;
      ++p;
      ws();
      if (p < end && *p == ']') { ++p; return v; }
      while (ok) {
        v.arr.push_back(val(depth + 1));
        ws();
        if (p < end && *p == ',') { ++p; continue; }
        if (p < end && *p == ']') { ++p; break; }
        ok = false;
      }
    } else if (*p == '"') {
      v.kind = JVal::STR;
      str(v.raw);
    } else if (lit("true")) {
      v.kind = JVal::BOOL; v.b = true; v.raw = "true";
    } else if (lit("false")) {
      v.kind = JVal::BOOL; v.raw = "false";
    } else if (lit("null")) {
      v.kind = JVal::NUL;
    } else {
      const char* s = p;
      while (p < end && (strchr("+-.eE", *p) || (*p >= '0' && *p <= '9'))) ++p;
      if (p == s) { ok = false; return v; }
      v.kind = JVal::NUM;
      v.raw.assign(s, p);
      v.num = strtod(v.raw.c_str(), nullptr);
    }
    return v;
  }
};

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
  Parser ps{json, json + strlen(json)};
  JVal root = ps.val();
  ps.ws();
  if (!ps.ok || ps.p != ps.end || root.kind != JVal::OBJ) return HB_ERR_INVALID;
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
