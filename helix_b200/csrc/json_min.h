// Minimal JSON reader (objects keep document order) shared by slot_config.cpp and tokenizer.cpp.
#pragma once
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <utility>
#include <vector>

namespace hbjson {

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


inline bool parse(const char* text, size_t len, JVal* out) {
  Parser ps{text, text + len};
  *out = ps.val();
  ps.ws();
  return ps.ok && ps.p == ps.end;
}

}  // namespace hbjson
