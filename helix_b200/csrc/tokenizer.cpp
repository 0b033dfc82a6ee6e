// Native byte-level BPE tokenizer + Llama-3 chat template (scope row F2): the backends the reference spawns tokenize
// inside the child process (vLLM loads the model's tokenizer.json; chat templates: Dockerfile.runner:61-74,188-191); an
// in-process runtime needs the same next to the engine, without Python.  Loads a HF `tokenizer.json` (BPE model, byte-level
// pre-tokenizer, added/special tokens) and reproduces `tokenizers` 0.22 bit for bit on the Llama-3 family's pipeline:
//   added tokens split out  ->  Split(regex) pre-tokenizer  ->  ByteLevel byte->char map  ->  BPE merges by rank
//   (ignore_merges: a pre-token that is itself in the vocabulary is taken whole).
// The Split regex of Llama-3 / GPT-4-style tokenizers is matched by a hand-written scanner (no regex engine):
//   (?i:'s|'t|'re|'ve|'m|'ll|'d) | [^\r\n\p{L}\p{N}]?\p{L}+ | \p{N}{1,3} | ?[^\s\p{L}\p{N}]+[\r\n]* | \s*[\r\n]+ | \s+(?!\S) | \s+
// Unicode classes come from unicode_tables.h (generated).  tests/test_tokenizer_cpu.py trains BPE vocabularies with the HF
// library on multilingual text and compares ids / decoded text on thousands of random strings.
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/helix_b200.h"
#include "json_min.h"
#include "unicode_tables.h"

namespace {

using hbjson::JVal;

bool in_ranges(const uint32_t (*r)[2], int n, uint32_t cp) {
  int lo = 0, hi = n - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    if (cp < r[mid][0]) hi = mid - 1;
    else if (cp > r[mid][1]) lo = mid + 1;
    else return true;
  }
  return false;
}
bool is_letter(uint32_t cp) { return cp < 128 ? ((cp | 32) >= 'a' && (cp | 32) <= 'z') : in_ranges(kUniLetter, kUniLetter_n, cp); }
bool is_number(uint32_t cp) { return cp < 128 ? (cp >= '0' && cp <= '9') : in_ranges(kUniNumber, kUniNumber_n, cp); }
bool is_space(uint32_t cp) {  // \s of the regex engine (Unicode White_Space)
  return (cp >= 9 && cp <= 13) || cp == 32 || cp == 0x85 || cp == 0xA0 || cp == 0x1680 || (cp >= 0x2000 && cp <= 0x200A) ||
         cp == 0x2028 || cp == 0x2029 || cp == 0x202F || cp == 0x205F || cp == 0x3000;
}
bool is_newline(uint32_t cp) { return cp == '\r' || cp == '\n'; }

// UTF-8 -> code points with byte offsets (invalid bytes become U+FFFD spanning one byte, like from_utf8_lossy would not —
// callers hand in valid UTF-8; this only keeps the scanner total)
struct Cp { uint32_t cp; uint32_t off; };
void decode_utf8(const std::string& s, std::vector<Cp>& out) {
  size_t i = 0;
  while (i < s.size()) {
    const unsigned char c = s[i];
    uint32_t cp = 0xFFFD;
    size_t n = 1;
    if (c < 0x80) { cp = c; }
    else if ((c >> 5) == 6 && i + 1 < s.size()) { cp = ((c & 0x1F) << 6) | (s[i + 1] & 0x3F); n = 2; }
    else if ((c >> 4) == 14 && i + 2 < s.size()) { cp = ((c & 0x0F) << 12) | ((s[i + 1] & 0x3F) << 6) | (s[i + 2] & 0x3F); n = 3; }
    else if ((c >> 3) == 30 && i + 3 < s.size()) {
      cp = ((c & 0x07) << 18) | ((s[i + 1] & 0x3F) << 12) | ((s[i + 2] & 0x3F) << 6) | (s[i + 3] & 0x3F);
      n = 4;
    }
    out.push_back(Cp{cp, (uint32_t)i});
    i += n;
  }
  out.push_back(Cp{0, (uint32_t)s.size()});  // sentinel (never matched: positions < size only)
}
void append_utf8(std::string& o, uint32_t cp) { hbjson::Parser::utf8(o, cp); }

// The Llama-3 Split pattern, leftmost alternative first; returns the end index (in code points) of the pre-token at i.
size_t scan_pretoken(const std::vector<Cp>& t, size_t i, size_t n) {
  auto cp = [&](size_t k) { return k < n ? t[k].cp : 0u; };
  auto lower = [](uint32_t c) { return (c >= 'A' && c <= 'Z') ? c + 32 : c; };
  // 1. contractions, case-insensitive: 's 't 're 've 'm 'll 'd
  if (cp(i) == '\'' && i + 1 < n) {
    const uint32_t a = lower(cp(i + 1)), b = i + 2 < n ? lower(cp(i + 2)) : 0;
    if (a == 's' || a == 't') return i + 2;
    if (a == 'r' && b == 'e') return i + 3;
    if (a == 'v' && b == 'e') return i + 3;
    if (a == 'm') return i + 2;
    if (a == 'l' && b == 'l') return i + 3;
    if (a == 'd') return i + 2;
  }
  // 2. [^\r\n\p{L}\p{N}]?\p{L}+
  {
    size_t k = i;
    const uint32_t c = cp(k);
    if (!is_newline(c) && !is_letter(c) && !is_number(c) && k + 1 < n && is_letter(cp(k + 1))) ++k;
    if (k < n && is_letter(cp(k))) {
      while (k < n && is_letter(cp(k))) ++k;
      return k;
    }
  }
  // 3. \p{N}{1,3}
  if (is_number(cp(i))) {
    size_t k = i;
    while (k < n && k < i + 3 && is_number(cp(k))) ++k;
    return k;
  }
  // 4.  ?[^\s\p{L}\p{N}]+[\r\n]*
  {
    size_t k = i;
    if (cp(k) == ' ' && k + 1 < n) ++k;
    auto other = [&](uint32_t c) { return !is_space(c) && !is_letter(c) && !is_number(c); };
    if (k < n && other(cp(k))) {
      while (k < n && other(cp(k))) ++k;
      while (k < n && is_newline(cp(k))) ++k;
      return k;
    }
  }
  // 5-7: whitespace runs
  if (is_space(cp(i))) {
    size_t e = i;
    while (e < n && is_space(cp(e))) ++e;
    // 5. \s*[\r\n]+ : greedy \s* backtracks to the LAST newline of the run
    for (size_t k = e; k > i; --k)
      if (is_newline(cp(k - 1))) return k;
    // 6. \s+(?!\S): the whole run at end of text, otherwise all but its last character (if anything is left)
    if (e == n) return e;
    if (e - i >= 2) return e - 1;
    // 7. \s+
    return e;
  }
  return i + 1;  // unreachable for valid input: every character class is covered above
}


// ---- BERT WordPiece pipeline (HF tokenizers BertNormalizer + BertPreTokenizer + WordPiece) ----
const uint32_t* seq_lookup(const uint32_t (*rows)[3], int n, const uint32_t* pool, uint32_t cp, int* len) {
  int lo = 0, hi = n - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    if (cp < rows[mid][0]) hi = mid - 1;
    else if (cp > rows[mid][0]) lo = mid + 1;
    else { *len = (int)rows[mid][2]; return pool + rows[mid][1]; }
  }
  *len = 0;
  return nullptr;
}
bool is_chinese(uint32_t c) {
  return (c >= 0x4E00 && c <= 0x9FFF) || (c >= 0x3400 && c <= 0x4DBF) || (c >= 0x20000 && c <= 0x2A6DF) || (c >= 0x2A700 && c <= 0x2B73F) ||
         (c >= 0x2B740 && c <= 0x2B81F) || (c >= 0x2B920 && c <= 0x2CEAF) || (c >= 0xF900 && c <= 0xFAFF) || (c >= 0x2F800 && c <= 0x2FA1F);
}
bool bert_is_whitespace(uint32_t c) { return c == '\t' || c == '\n' || c == '\r' || is_space(c); }
bool bert_is_control(uint32_t c) {
  if (c == '\t' || c == '\n' || c == '\r') return false;
  return c < 128 ? (c < 32 || c == 127) : in_ranges(kUniOther, kUniOther_n, c);
}
bool bert_is_punct(uint32_t c) {
  if (c < 128) return (c >= 33 && c <= 47) || (c >= 58 && c <= 64) || (c >= 91 && c <= 96) || (c >= 123 && c <= 126);
  return in_ranges(kUniPunct, kUniPunct_n, c);
}
void nfd_append(std::vector<uint32_t>& out, uint32_t c) {
  if (c >= 0xAC00 && c <= 0xD7A3) {  // Hangul syllable: algorithmic decomposition
    const uint32_t s = c - 0xAC00, l = 0x1100 + s / 588, v = 0x1161 + (s % 588) / 28, t = 0x11A7 + s % 28;
    out.push_back(l);
    out.push_back(v);
    if (t != 0x11A7) out.push_back(t);
    return;
  }
  int n = 0;
  const uint32_t* d = c < 0xC0 ? nullptr : seq_lookup(kUniNfd, kUniNfd_n, kUniNfd_pool, c, &n);
  if (d) out.insert(out.end(), d, d + n); else out.push_back(c);
}
// BertNormalizer(clean_text, handle_chinese_chars, strip_accents, lowercase) in the library's order
std::vector<uint32_t> bert_normalize(const std::vector<Cp>& t, size_t n, bool clean, bool chinese, bool strip, bool lower) {
  std::vector<uint32_t> a, b;
  a.reserve(n + 8);
  for (size_t i = 0; i < n; ++i) {
    const uint32_t c = t[i].cp;
    if (clean) {
      if (c == 0 || c == 0xFFFD || bert_is_control(c)) continue;
      a.push_back(bert_is_whitespace(c) ? ' ' : c);
    } else {
      a.push_back(c);
    }
  }
  if (chinese) {
    b.clear();
    for (uint32_t c : a) {
      if (is_chinese(c)) { b.push_back(' '); b.push_back(c); b.push_back(' '); } else b.push_back(c);
    }
    a.swap(b);
  }
  if (strip) {
    b.clear();
    for (uint32_t c : a) nfd_append(b, c);
    a.clear();
    for (uint32_t c : b)
      if (!(c >= 0x300 && in_ranges(kUniMarkNonspacing, kUniMarkNonspacing_n, c))) a.push_back(c);
  }
  if (lower) {
    b.clear();
    for (uint32_t c : a) {
      if (c < 128) { b.push_back((c >= 'A' && c <= 'Z') ? c + 32 : c); continue; }
      int ln = 0;
      const uint32_t* d = seq_lookup(kUniLower, kUniLower_n, kUniLower_pool, c, &ln);
      if (d) b.insert(b.end(), d, d + ln); else b.push_back(c);
    }
    a.swap(b);
  }
  return a;
}

}  // namespace

struct hb_tokenizer {
  std::unordered_map<std::string, int32_t> vocab;
  std::vector<std::string> id_to_token;
  std::unordered_map<std::string, int32_t> merge_rank;  // "left right" -> rank
  struct Added { std::string content; int32_t id; bool special; };
  std::vector<Added> added;  // longest content first
  bool ignore_merges = false;
  // WordPiece (BERT-family encoders)
  bool wordpiece = false, bn_clean = true, bn_chinese = true, bn_strip = true, bn_lower = true;
  std::string wp_prefix = "##", unk_token = "[UNK]";
  int32_t unk_id = -1, cls_id = -1, sep_id = -1;
  size_t wp_max_chars = 100;
  std::string byte_char[256];                       // ByteLevel: byte -> UTF-8 of its stand-in character
  std::unordered_map<uint32_t, uint8_t> char_byte;  // stand-in code point -> byte
  std::string error;

  void init_bytes() {
    // GPT-2 byte-level alphabet: printable latin-1 bytes map to themselves, the rest to U+0100.. in order
    int extra = 0;
    for (int b = 0; b < 256; ++b) {
      const bool keep = (b >= 33 && b <= 126) || (b >= 161 && b <= 172) || (b >= 174 && b <= 255);
      const uint32_t cp = keep ? (uint32_t)b : 256u + (uint32_t)extra++;
      byte_char[b].clear();
      append_utf8(byte_char[b], cp);
      char_byte[cp] = (uint8_t)b;
    }
  }

  bool load(const std::string& json_text) {
    JVal root;
    if (!hbjson::parse(json_text.data(), json_text.size(), &root) || root.kind != JVal::OBJ) { error = "tokenizer.json: not a JSON object"; return false; }
    const JVal* model = root.get("model");
    const JVal* type = model ? model->get("type") : nullptr;
    if (!model || model->kind != JVal::OBJ) { error = "tokenizer.json: model missing"; return false; }
    if (type && type->kind == JVal::STR && type->raw == "WordPiece") wordpiece = true;
    else if (type && type->kind == JVal::STR && type->raw != "BPE") { error = "tokenizer.json: model.type must be BPE or WordPiece"; return false; }
    const JVal* v = model->get("vocab");
    if (!v || v->kind != JVal::OBJ) { error = "tokenizer.json: model.vocab missing"; return false; }
    size_t max_id = 0;
    for (const auto& kv : v->obj) {
      if (kv.second.kind != JVal::NUM) continue;
      const int32_t id = (int32_t)kv.second.num;
      vocab[kv.first] = id;
      max_id = std::max<size_t>(max_id, (size_t)id);
    }
    if (const JVal* m = model->get("merges"); m && m->kind == JVal::ARR) {
      int32_t rank = 0;
      for (const JVal& e : m->arr) {
        if (e.kind == JVal::STR) merge_rank[e.raw] = rank++;                        // "left right"
        else if (e.kind == JVal::ARR && e.arr.size() == 2) merge_rank[e.arr[0].raw + " " + e.arr[1].raw] = rank++;  // ["left","right"]
      }
    }
    if (const JVal* im = model->get("ignore_merges"); im && im->kind == JVal::BOOL) ignore_merges = im->b;
    if (const JVal* at = root.get("added_tokens"); at && at->kind == JVal::ARR) {
      for (const JVal& a : at->arr) {
        const JVal* c = a.get("content");
        const JVal* id = a.get("id");
        const JVal* sp = a.get("special");
        if (!c || !id || c->kind != JVal::STR || id->kind != JVal::NUM) continue;
        added.push_back(Added{c->raw, (int32_t)id->num, sp && sp->kind == JVal::BOOL && sp->b});
        max_id = std::max<size_t>(max_id, (size_t)id->num);
      }
      std::stable_sort(added.begin(), added.end(), [](const Added& x, const Added& y) { return x.content.size() > y.content.size(); });
    }
    id_to_token.assign(max_id + 1, std::string());
    for (const auto& kv : vocab) id_to_token[kv.second] = kv.first;
    for (const Added& a : added) id_to_token[a.id] = a.content;
    init_bytes();
    if (wordpiece) {
      if (const JVal* p = model->get("continuing_subword_prefix"); p && p->kind == JVal::STR) wp_prefix = p->raw;
      if (const JVal* u = model->get("unk_token"); u && u->kind == JVal::STR) unk_token = u->raw;
      if (const JVal* mx = model->get("max_input_chars_per_word"); mx && mx->kind == JVal::NUM) wp_max_chars = (size_t)mx->num;
      unk_id = added_id(unk_token);
      cls_id = added_id("[CLS]");
      sep_id = added_id("[SEP]");
      if (const JVal* nz = root.get("normalizer"); nz && nz->kind == JVal::OBJ) {
        auto flag = [&](const char* k, bool dflt) {
          const JVal* v = nz->get(k);
          return (v && v->kind == JVal::BOOL) ? v->b : dflt;
        };
        bn_clean = flag("clean_text", true);
        bn_chinese = flag("handle_chinese_chars", true);
        bn_lower = flag("lowercase", true);
        const JVal* sa = nz->get("strip_accents");
        bn_strip = (sa && sa->kind == JVal::BOOL) ? sa->b : bn_lower;  // null: follows lowercase
      }
    }
    return true;
  }

  int32_t added_id(const std::string& content) const {
    for (const Added& a : added)
      if (a.content == content) return a.id;
    auto it = vocab.find(content);
    return it == vocab.end() ? -1 : it->second;
  }

  // BPE over one pre-token given as its byte-level string
  void bpe(const std::string& word, std::vector<int32_t>& out) const {
    if (ignore_merges) {
      auto it = vocab.find(word);
      if (it != vocab.end()) { out.push_back(it->second); return; }
    }
    std::vector<std::string> sym;  // one entry per byte-level character
    for (size_t i = 0; i < word.size();) {
      size_t n = 1;
      const unsigned char c = word[i];
      if ((c >> 5) == 6) n = 2; else if ((c >> 4) == 14) n = 3; else if ((c >> 3) == 30) n = 4;
      sym.emplace_back(word, i, n);
      i += n;
    }
    while (sym.size() > 1) {
      int32_t best = INT32_MAX;
      size_t at = 0;
      for (size_t i = 0; i + 1 < sym.size(); ++i) {
        auto it = merge_rank.find(sym[i] + " " + sym[i + 1]);
        if (it != merge_rank.end() && it->second < best) { best = it->second; at = i; }
      }
      if (best == INT32_MAX) break;
      sym[at] += sym[at + 1];
      sym.erase(sym.begin() + at + 1);
    }
    for (const std::string& s : sym) {
      auto it = vocab.find(s);
      if (it != vocab.end()) {
        out.push_back(it->second);
      } else {  // cannot happen with a byte-level vocabulary that holds all 256 byte characters
        for (size_t i = 0; i < s.size();) {
          size_t n = 1;
          const unsigned char c = s[i];
          if ((c >> 5) == 6) n = 2; else if ((c >> 4) == 14) n = 3; else if ((c >> 3) == 30) n = 4;
          auto b = vocab.find(s.substr(i, n));
          if (b != vocab.end()) out.push_back(b->second);
          i += n;
        }
      }
    }
  }

  // WordPiece: normalise, split on whitespace / punctuation, greedy longest-match-first with the continuing prefix
  void encode_wordpiece(const std::string& text, std::vector<int32_t>& out) const {
    std::vector<Cp> t;
    decode_utf8(text, t);
    const std::vector<uint32_t> nrm = bert_normalize(t, t.size() - 1, bn_clean, bn_chinese, bn_strip, bn_lower);
    std::vector<std::string> chars;  // the current word, one UTF-8 string per code point
    auto flush = [&]() {
      if (chars.empty()) return;
      if (chars.size() > wp_max_chars) { out.push_back(unk_id); chars.clear(); return; }
      std::vector<int32_t> pieces;
      size_t start = 0;
      bool bad = false;
      while (start < chars.size()) {
        size_t end = chars.size();
        int32_t found = -1;
        for (; end > start; --end) {
          std::string sub = start ? wp_prefix : std::string();
          for (size_t k = start; k < end; ++k) sub += chars[k];
          auto it = vocab.find(sub);
          if (it != vocab.end()) { found = it->second; break; }
        }
        if (found < 0) { bad = true; break; }
        pieces.push_back(found);
        start = end;
      }
      if (bad) out.push_back(unk_id); else out.insert(out.end(), pieces.begin(), pieces.end());
      chars.clear();
    };
    for (uint32_t c : nrm) {
      if (bert_is_whitespace(c)) { flush(); continue; }
      std::string u;
      append_utf8(u, c);
      if (bert_is_punct(c)) { flush(); chars.push_back(u); flush(); continue; }
      chars.push_back(u);
    }
    flush();
  }

  void encode_plain(const std::string& text, std::vector<int32_t>& out) const {
    if (text.empty()) return;
    if (wordpiece) { encode_wordpiece(text, out); return; }
    std::vector<Cp> t;
    decode_utf8(text, t);
    const size_t n = t.size() - 1;
    std::string word;
    for (size_t i = 0; i < n;) {
      const size_t e = scan_pretoken(t, i, n);
      word.clear();
      for (uint32_t b = t[i].off; b < t[e].off; ++b) word += byte_char[(unsigned char)text[b]];
      bpe(word, out);
      i = e;
    }
  }

  void encode(const std::string& text, bool parse_special, std::vector<int32_t>& out) const {
    if (added.empty()) { encode_plain(text, out); return; }
    size_t start = 0, i = 0;
    while (i < text.size()) {
      const Added* hit = nullptr;
      for (const Added& a : added) {  // longest first
        if (a.special && !parse_special) continue;
        if (!a.content.empty() && text.compare(i, a.content.size(), a.content) == 0) { hit = &a; break; }
      }
      if (hit) {
        encode_plain(text.substr(start, i - start), out);
        out.push_back(hit->id);
        i += hit->content.size();
        start = i;
      } else {
        ++i;
      }
    }
    encode_plain(text.substr(start), out);
  }

  std::string decode(const int32_t* ids, int n, bool skip_special) const {
    if (wordpiece) {  // WordPiece decoder: pieces joined by spaces, continuing pieces glued on (cleanup off)
      std::string s;
      bool first = true;
      for (int k = 0; k < n; ++k) {
        const int32_t id = ids[k];
        if (id < 0 || (size_t)id >= id_to_token.size()) continue;
        bool special = false;
        for (const Added& a : added)
          if (a.id == id) { special = a.special; break; }
        if (special && skip_special) continue;
        const std::string& tok = id_to_token[id];
        if (!first && tok.compare(0, wp_prefix.size(), wp_prefix) == 0) { s += tok.substr(wp_prefix.size()); }
        else { if (!first) s += ' '; s += tok; }
        first = false;
      }
      return s;
    }
    std::string bytes;
    for (int k = 0; k < n; ++k) {
      const int32_t id = ids[k];
      if (id < 0 || (size_t)id >= id_to_token.size()) continue;
      bool is_added = false, special = false;
      for (const Added& a : added)
        if (a.id == id) { is_added = true; special = a.special; break; }
      if (is_added) {
        if (!(special && skip_special)) bytes += id_to_token[id];  // added tokens are literal text, not byte-level
        continue;
      }
      const std::string& tok = id_to_token[id];
      std::vector<Cp> t;
      decode_utf8(tok, t);
      for (size_t i = 0; i + 1 < t.size(); ++i) {
        auto it = char_byte.find(t[i].cp);
        if (it != char_byte.end()) bytes += (char)it->second;
      }
    }
    return bytes;  // UTF-8 of whole characters when the ids form whole characters; the caller handles partial tails
  }
};

extern "C" {

int hb_tok_load(const char* tokenizer_json_path, hb_tokenizer** out) {
  if (!tokenizer_json_path || !out) return HB_ERR_INVALID;
  *out = nullptr;
  FILE* f = fopen(tokenizer_json_path, "rb");
  if (!f) return HB_ERR_NOT_FOUND;
  std::string text;
  char buf[1 << 16];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, n);
  fclose(f);
  auto t = std::make_unique<hb_tokenizer>();
  if (!t->load(text)) return HB_ERR_INVALID;
  *out = t.release();
  return HB_OK;
}
void hb_tok_free(hb_tokenizer* t) { delete t; }
int32_t hb_tok_vocab_size(hb_tokenizer* t) { return t ? (int32_t)t->id_to_token.size() : 0; }
int32_t hb_tok_token_id(hb_tokenizer* t, const char* token) { return (t && token) ? t->added_id(token) : -1; }

int hb_tok_encode(hb_tokenizer* t, const char* utf8, int32_t parse_special, int32_t* out, int32_t cap, int32_t* n) {
  if (!t || !utf8 || !n) return HB_ERR_INVALID;
  std::vector<int32_t> ids;
  if (parse_special == 2) {  // the model's own framing, what HF's encode(add_special_tokens=True) adds
    if (t->wordpiece && t->cls_id >= 0) ids.push_back(t->cls_id);
    if (!t->wordpiece) {
      const int32_t bot = t->added_id("<|begin_of_text|>");
      if (bot >= 0) ids.push_back(bot);
    }
  }
  t->encode(utf8, parse_special != 0, ids);
  if (parse_special == 2 && t->wordpiece && t->sep_id >= 0) ids.push_back(t->sep_id);
  *n = (int32_t)ids.size();
  if ((int32_t)ids.size() > cap || (!out && !ids.empty())) return HB_ERR_BUSY;  // *n tells the needed capacity
  if (!ids.empty()) memcpy(out, ids.data(), ids.size() * 4);
  return HB_OK;
}

int hb_tok_decode(hb_tokenizer* t, const int32_t* ids, int32_t n_ids, int32_t skip_special, char* out, size_t cap, size_t* len) {
  if (!t || (n_ids > 0 && !ids) || !len) return HB_ERR_INVALID;
  const std::string s = t->decode(ids, n_ids, skip_special != 0);
  *len = s.size();
  if (s.size() + 1 > cap || !out) return HB_ERR_BUSY;  // *len tells the needed capacity (plus the terminator)
  memcpy(out, s.data(), s.size());
  out[s.size()] = 0;
  return HB_OK;
}

// Llama-3 instruct chat template: <|begin_of_text|> then per message
// <|start_header_id|>role<|end_header_id|>\n\ncontent<|eot_id|>, then the assistant header to generate after.
int hb_tok_chat_llama3(hb_tokenizer* t, const char* const* roles, const char* const* contents, int32_t n_msgs, int32_t* out,
                       int32_t cap, int32_t* n) {
  if (!t || !n || (n_msgs > 0 && (!roles || !contents))) return HB_ERR_INVALID;
  const int32_t bot = t->added_id("<|begin_of_text|>"), sh = t->added_id("<|start_header_id|>"),
                eh = t->added_id("<|end_header_id|>"), eot = t->added_id("<|eot_id|>");
  if (bot < 0 || sh < 0 || eh < 0 || eot < 0) return HB_ERR_NOT_FOUND;  // not a Llama-3 style vocabulary
  (void)sh; (void)eh; (void)eot;
  // the template is rendered as ONE string and encoded with special tokens parsed, exactly as HF's apply_chat_template +
  // tokenizer call does (pre-token boundaries across "\n\n" + content depend on it)
  std::string text = "<|begin_of_text|>";
  for (int i = 0; i < n_msgs; ++i) {
    text += "<|start_header_id|>";
    text += roles[i] ? roles[i] : "user";
    text += "<|end_header_id|>\n\n";
    text += contents[i] ? contents[i] : "";
    text += "<|eot_id|>";
  }
  text += "<|start_header_id|>assistant<|end_header_id|>\n\n";
  std::vector<int32_t> ids;
  t->encode(text, true, ids);
  *n = (int32_t)ids.size();
  if ((int32_t)ids.size() > cap || !out) return HB_ERR_BUSY;
  memcpy(out, ids.data(), ids.size() * 4);
  return HB_OK;
}

}  // extern "C"
