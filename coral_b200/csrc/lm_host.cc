// ARPA parser and table builder (host). See lm_host.h.
#include "lm_host.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

namespace coral {

bool utf8_to_u32(const std::string& s, std::u32string& out) {
  out.clear();
  size_t i = 0, n = s.size();
  while (i < n) {
    unsigned char c = (unsigned char)s[i];
    uint32_t cp;
    int extra;
    if (c < 0x80) { cp = c; extra = 0; }
    else if ((c >> 5) == 0x6) { cp = c & 0x1F; extra = 1; }
    else if ((c >> 4) == 0xE) { cp = c & 0x0F; extra = 2; }
    else if ((c >> 3) == 0x1E) { cp = c & 0x07; extra = 3; }
    else return false;
    for (int k = 1; k <= extra; ++k) {
      if (i + k >= n) return false;
      unsigned char cc = (unsigned char)s[i + k];
      if ((cc >> 6) != 0x2) return false;
      cp = (cp << 6) | (cc & 0x3F);
    }
    out.push_back((char32_t)cp);
    i += 1 + extra;
  }
  return true;
}

uint64_t hash_word(const std::u32string& w) {
  uint64_t h = kWordHashSeed;
  for (char32_t c : w) h = word_hash_push(h, (uint32_t)c);
  return h;
}

static uint64_t next_pow2(uint64_t x) {
  uint64_t p = 1;
  while (p < x) p <<= 1;
  return p;
}

namespace {
struct Pending {
  uint64_t key;
  float prob, backoff;
  int n;
};
}  // namespace

int load_arpa(const char* path, HostLm& lm, std::string& err) {
  FILE* f = fopen(path, "rb");
  if (!f) { err = std::string("cannot open ARPA file: ") + path; return -2; }
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::string buf;
  buf.resize((size_t)sz);
  if (sz > 0 && fread(&buf[0], 1, (size_t)sz, f) != (size_t)sz) {
    fclose(f);
    err = "short read on ARPA file";
    return -2;
  }
  fclose(f);

  lm = HostLm();
  lm.path = path;
  lm.words.push_back("<unk>");
  lm.vocab.emplace("<unk>", 0u);
  lm.uni.push_back(UniEntry{-100.0f, 0.0f});  // kenlm unknown_missing_logprob
  bool saw_unk = false;
  std::vector<uint64_t> declared(kMaxOrder + 2, 0);
  int max_declared = 0;
  std::vector<Pending> pending;
  int section = 0;
  bool ended = false;

  size_t pos = 0, n = buf.size();
  std::vector<uint32_t> ids;
  while (pos < n && !ended) {
    size_t eol = buf.find('\n', pos);
    if (eol == std::string::npos) eol = n;
    size_t b = pos, e = eol;
    pos = eol + 1;
    while (e > b && (buf[e - 1] == '\r')) --e;
    // blank line?
    size_t k = b;
    while (k < e && (buf[k] == ' ' || buf[k] == '\t')) ++k;
    if (k == e) continue;
    if (buf[b] == '\\') {
      std::string tag(buf, b, e - b);
      while (!tag.empty() && (tag.back() == ' ' || tag.back() == '\t')) tag.pop_back();
      if (tag == "\\data\\") section = 0;
      else if (tag == "\\end\\") ended = true;
      else {
        size_t dash = tag.find('-');
        if (dash == std::string::npos || tag.size() < 8 || tag.substr(dash) != "-grams:") {
          err = "unknown ARPA section " + tag;
          return -2;
        }
        section = atoi(tag.c_str() + 1);
        if (section < 1 || section > kMaxOrder) {
          err = "ARPA order above the supported maximum (6): " + tag;
          return -2;
        }
      }
      continue;
    }
    if (section == 0) {
      if (e - b > 6 && !strncmp(&buf[b], "ngram ", 6)) {
        int nn = atoi(&buf[b + 6]);
        const char* eq = (const char*)memchr(&buf[b], '=', e - b);
        if (eq && nn >= 1 && nn <= kMaxOrder) {
          declared[nn] = strtoull(eq + 1, nullptr, 10);
          max_declared = std::max(max_declared, nn);
        } else if (nn > kMaxOrder) {
          err = "ARPA order above the supported maximum (6)";
          return -2;
        }
      }
      continue;
    }
    // "<prob>\t<w1 w2 ...>[\t<backoff>]"
    const char* line = &buf[b];
    const char* lend = &buf[e];
    const char* t1 = (const char*)memchr(line, '\t', lend - line);
    if (!t1) { err = "malformed ARPA line (no tab): " + std::string(line, lend - line); return -2; }
    float prob = (float)strtod(line, nullptr);  // double rounding like float32(float(str))
    const char* wbeg = t1 + 1;
    const char* t2 = (const char*)memchr(wbeg, '\t', lend - wbeg);
    const char* wend = t2 ? t2 : lend;
    float backoff = 0.0f;
    if (t2) backoff = (float)strtod(t2 + 1, nullptr);
    if (section == 1) {
      std::string w(wbeg, wend - wbeg);
      if (w.find(' ') != std::string::npos) { err = "expected 1 word in: " + std::string(line, lend - line); return -2; }
      if (w == "<unk>") {
        if (!saw_unk) { saw_unk = true; lm.uni[0] = UniEntry{prob, backoff}; }
      } else if (lm.vocab.find(w) == lm.vocab.end()) {
        lm.vocab.emplace(w, (uint32_t)lm.words.size());
        lm.words.push_back(w);
        lm.uni.push_back(UniEntry{prob, backoff});
      }  // else: duplicate unigram, first occurrence keeps the slot
    } else {
      ids.clear();
      const char* p = wbeg;
      while (p <= wend) {
        const char* sp = (const char*)memchr(p, ' ', wend - p);
        const char* we = sp ? sp : wend;
        auto it = lm.vocab.find(std::string(p, we - p));
        ids.push_back(it == lm.vocab.end() ? 0u : it->second);
        if (!sp) break;
        p = sp + 1;
      }
      if ((int)ids.size() != section) {
        err = "wrong word count in " + std::to_string(section) + "-gram line: " + std::string(line, lend - line);
        return -2;
      }
      // chain key: predicted word first, then context most-recent-first
      uint64_t key = kNgSeed;
      for (int i = section - 1; i >= 0; --i) key = ng_key_push(key, ids[i]);
      pending.push_back(Pending{key, prob, backoff, section});
    }
  }
  int max_seen = 1;
  for (const auto& p : pending) max_seen = std::max(max_seen, p.n);
  lm.order = max_declared ? max_declared : max_seen;
  if (max_seen > lm.order) lm.order = max_seen;
  lm.counts.assign(declared.begin() + 1, declared.begin() + 1 + lm.order);
  lm.loaded.assign(lm.order, 0);
  lm.loaded[0] = lm.uni.size();
  {
    auto it = lm.vocab.find("<s>");
    lm.bos_id = it == lm.vocab.end() ? 0u : it->second;
    it = lm.vocab.find("</s>");
    lm.eos_id = it == lm.vocab.end() ? 0u : it->second;
  }
  if (!pending.empty()) {
    uint64_t cap = next_pow2(std::max<uint64_t>(16, pending.size() * 2));
    lm.ng.assign(cap, NgSlot{0, 0.0f, 0.0f});
    lm.ng_mask = cap - 1;
    for (const auto& p : pending) {
      uint64_t i = (p.key >> 20) & lm.ng_mask;
      for (;;) {
        if (lm.ng[i].key == 0) { lm.ng[i] = NgSlot{p.key, p.prob, p.backoff}; break; }
        if (lm.ng[i].key == p.key) {
          err = "duplicate n-gram in ARPA file (or 64-bit chain-key collision)";
          return -2;
        }
        i = (i + 1) & lm.ng_mask;
      }
      lm.loaded[p.n - 1]++;
    }
  } else {
    lm.ng.assign(16, NgSlot{0, 0.0f, 0.0f});
    lm.ng_mask = 15;
  }
  return 0;
}

int build_lexicon(const HostLm& lm, const std::vector<std::u32string>* unigrams, HostLexicon& out,
                  std::string& err) {
  struct Info { uint32_t wid, flags; std::u32string s; };
  std::unordered_map<uint64_t, Info> map;
  map.reserve(lm.words.size() * 6);
  auto add = [&](const std::u32string& w, uint32_t wid, uint32_t full_flags, uint32_t prefix_flags) -> bool {
    uint64_t h = kWordHashSeed;
    std::u32string pre;
    for (size_t i = 0; i < w.size(); ++i) {
      h = word_hash_push(h, (uint32_t)w[i]);
      pre.push_back(w[i]);
      auto it = map.find(h);
      const bool full = (i + 1 == w.size());
      if (it == map.end()) {
        it = map.emplace(h, Info{0u, 0u, pre}).first;
      } else if (it->second.s != pre) {
        return false;  // two distinct strings share a 64-bit hash
      }
      it->second.flags |= prefix_flags;
      if (full) {
        it->second.flags |= full_flags;
        if (full_flags & kLexInLm) it->second.wid = wid;
      }
    }
    return true;
  };
  std::u32string u;
  for (uint32_t id = 1; id < lm.words.size(); ++id) {
    if (!utf8_to_u32(lm.words[id], u)) { err = "invalid UTF-8 in LM vocabulary"; return -2; }
    if (!add(u, id, kLexInLm, 0)) { err = "word-hash collision in lexicon"; return -2; }
  }
  uint64_t n_uni = 0;
  if (unigrams) {
    for (const auto& w : *unigrams) {
      // pyctcdecode keeps only unigrams that are in the kenlm vocabulary
      if (w.empty()) continue;
      uint64_t h = hash_word(w);
      auto it = map.find(h);
      if (it == map.end() || it->second.s != w || !(it->second.flags & kLexInLm)) continue;
      if (!add(w, 0, kLexInUnigrams, kLexPrefixOfUnigram)) { err = "word-hash collision in lexicon"; return -2; }
      ++n_uni;
    }
  }
  out.has_unigrams = n_uni > 0 ? 1 : 0;
  out.n_entries = map.size();
  uint64_t cap = next_pow2(std::max<uint64_t>(16, map.size() * 2));
  out.lex.assign(cap, LexSlot{0, 0, 0});
  out.lex_mask = cap - 1;
  for (const auto& kv : map) {
    uint64_t i = (kv.first >> 20) & out.lex_mask;
    while (out.lex[i].key != 0) i = (i + 1) & out.lex_mask;
    out.lex[i] = LexSlot{kv.first, kv.second.wid, kv.second.flags};
  }
  return 0;
}

}  // namespace coral
