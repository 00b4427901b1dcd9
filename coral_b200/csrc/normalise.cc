// C-ABI (host only): the text normaliser that sits between decoding and scoring.
//
// Replaces the text part of coral.data.process_example (R:src/coral/data.py:616-701, called with
// audio_column=None at R:src/coral/evaluate.py:61-72 and R:src/coral/validation.py:121-132) and
// coral.utils.convert_numeral_to_words (R:src/coral/utils.py:303-472, NUMERAL_REGEX at :31). With
// decoding at a few milliseconds per thousand utterances, the reference's per-utterance Python
// (regex, NFKC, dict replaces) becomes the serial tail of evaluate(); here a batch of transcripts is
// normalised by a few host threads.
//
// The steps and their order are the reference's:
//   1. numerals -> Danish words (re.split on NUMERAL_REGEX, each piece through
//      convert_numeral_to_words)                                        [convert_numerals]
//   2. str.lower()                                                       [lower_case]
//   3. FILLER_WORDS_PATTERN.sub("")   \b(eh+m*|øh+m*|h+m+|m+h+)\b, IGNORECASE
//   4. unicodedata.normalize("NFKC")
//   5. for key, value in conversion_dict.items(): doc = doc.replace(key, value)   (in order)
//   6. [^<characters_to_keep + ' |'>] (IGNORECASE) -> " " on doc.strip()  [characters_to_keep]
//   7. " +" -> " "
//   8. strip every line, strip leading / trailing newlines
// CPython's Unicode behaviour (lower, NFKC, \b, \d, str.strip, re's case folding) comes from tables
// dumped from the interpreter itself (gen_unicode_tables.py -> build/unicode_tables.h).
// What is NOT restated is refused per string (status 1) instead of being approximated: GREEK CAPITAL
// SIGMA under lower_case (CPython's final-sigma rule) and non-ASCII decimal digits under
// convert_numerals (the reference itself raises KeyError on them).
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <unordered_set>
#include <utility>
#include <vector>

#include "../../include/coral_b200.h"
#include "unicode_tables.h"

namespace coral {
int32_t fail(int32_t code, const std::string& msg);

namespace {

using U32 = std::u32string;

template <class K>
int find_key(const K* keys, int n, K k) {
  const K* e = keys + n;
  const K* p = std::lower_bound(keys, e, k);
  return (p != e && *p == k) ? (int)(p - keys) : -1;
}
bool in_ranges(const uint32_t* lo, const uint32_t* hi, int n, uint32_t c) {
  const uint32_t* p = std::upper_bound(lo, lo + n, c);
  if (p == lo) return false;
  return c <= hi[(p - lo) - 1];
}
bool is_word(uint32_t c) {  // re's \w for str patterns: alphanumeric or underscore
  if (c < 128) return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_';
  if (c >= 0xC0 && c < 0x100) return c != 0xD7 && c != 0xF7;  // Latin-1 letters (the Danish ones live here)
  return in_ranges(uni::kWordLo, uni::kWordHi, uni::kWordN, c);
}
bool is_space(uint32_t c) {
  if (c < 128) return (c >= 9 && c <= 13) || (c >= 28 && c <= 32);
  return in_ranges(uni::kSpaceLo, uni::kSpaceHi, uni::kSpaceN, c);
}
bool is_nonascii_digit(uint32_t c) { return c >= 128 && in_ranges(uni::kDigitLo, uni::kDigitHi, uni::kDigitN, c); }
bool is_ascii_digit(uint32_t c) { return c >= '0' && c <= '9'; }
uint32_t lower_simple(uint32_t c) {
  if (c < 128) return (c >= 'A' && c <= 'Z') ? c + 32 : c;
  if (c >= 0xE0 && c < 0x100) return c;  // Latin-1 lower-case letters
  const int i = find_key(uni::kLowSimpleKey, uni::kLowSimpleN, c);
  return i < 0 ? c : uni::kLowSimpleVal[i];
}
uint8_t ccc(uint32_t c) {
  if (c < 0x300) return 0;
  const int i = find_key(uni::kCccKey, uni::kCccN, c);
  return i < 0 ? 0 : uni::kCccVal[i];
}

// ---------------------------------------------------------------- NFKC (UAX #15)
constexpr uint32_t SB = 0xAC00, LB = 0x1100, VB = 0x1161, TB = 0x11A7, LC = 19, VC = 21, TC = 28, NC = VC * TC, SC = LC * NC;

uint32_t compose_pair(uint32_t a, uint32_t b) {
  if (a >= LB && a < LB + LC && b >= VB && b < VB + VC) return SB + ((a - LB) * VC + (b - VB)) * TC;
  if (a >= SB && a < SB + SC && (a - SB) % TC == 0 && b > TB && b < TB + TC) return a + (b - TB);
  const int i = find_key(uni::kCompKey, uni::kCompN, ((uint64_t)a << 32) | b);
  return i < 0 ? 0 : uni::kCompVal[i];
}

void nfkc(U32& s) {
  bool plain = true;  // ASCII and Latin-1 letters without a compatibility mapping are the common case
  for (char32_t c : s)
    if (c >= 0xA0 && !(c >= 0xC0 && c < 0x132)) { plain = false; break; }
  if (plain) return;
  U32 d;
  d.reserve(s.size() + 8);
  for (char32_t c : s) {
    if (c >= SB && c < SB + SC) {
      const uint32_t si = c - SB;
      d.push_back(LB + si / NC);
      d.push_back(VB + (si % NC) / TC);
      if (si % TC) d.push_back(TB + si % TC);
      continue;
    }
    const int i = c < 0xA0 ? -1 : find_key(uni::kNfkdKey, uni::kNfkdN, (uint32_t)c);
    if (i < 0) d.push_back(c);
    else for (uint32_t k = uni::kNfkdOff[i]; k < uni::kNfkdOff[i + 1]; ++k) d.push_back(uni::kNfkdData[k]);
  }
  // canonical ordering: stable sort of every run of non-starters by combining class
  for (size_t i = 0; i < d.size();) {
    if (ccc(d[i]) == 0) { ++i; continue; }
    size_t j = i;
    while (j < d.size() && ccc(d[j]) != 0) ++j;
    std::stable_sort(d.begin() + i, d.begin() + j, [](char32_t x, char32_t y) { return ccc(x) < ccc(y); });
    i = j;
  }
  if (d.empty()) { s.clear(); return; }
  // canonical composition
  size_t starter_pos = 0, comp_pos = 1;
  uint32_t starter = d[0];
  int last_class = ccc(starter);
  if (last_class != 0) last_class = 256;  // a string that opens with a combining mark has no starter yet
  for (size_t k = 1; k < d.size(); ++k) {
    const uint32_t ch = d[k];
    const int cc = ccc(ch);
    const uint32_t comp = compose_pair(starter, ch);
    if (comp != 0 && (last_class < cc || last_class == 0)) {
      d[starter_pos] = comp;
      starter = comp;
    } else {
      if (cc == 0) { starter_pos = comp_pos; starter = ch; }
      last_class = cc;
      d[comp_pos++] = ch;
    }
  }
  d.resize(comp_pos);
  s.swap(d);
}

// ------------------------------------------------------- numerals (R:src/coral/utils.py:303-472)
std::string replace_all(std::string s, const std::string& a, const std::string& b) {
  size_t pos = 0;
  while ((pos = s.find(a, pos)) != std::string::npos) { s.replace(pos, a.size(), b); pos += b.size(); }
  return s;
}
std::string squeeze_strip(const std::string& s) {  // re.sub(r" +", " ", s).strip()
  std::string o;
  for (char c : s) if (!(c == ' ' && !o.empty() && o.back() == ' ')) o.push_back(c);
  size_t a = 0, b = o.size();
  while (a < b && is_space((unsigned char)o[a])) ++a;
  while (b > a && is_space((unsigned char)o[b - 1])) --b;
  return o.substr(a, b - a);
}
std::string lstrip0(const std::string& s) { size_t a = 0; while (a < s.size() && s[a] == '0') ++a; return s.substr(a); }
size_t int_len(const std::string& s) { const std::string t = lstrip0(s); return t.empty() ? 1 : t.size(); }  // len(str(int(s)))

// `n`: ASCII digits without separators (what the reference holds after numeral.replace(".", "")),
// already known to match NUMERAL_REGEX ("" or a leading zero falls out of the regex: returned as is)
std::string number_words(const std::string& n, bool inside) {
  if (n.empty() || (n[0] == '0' && n.size() > 1)) return n;
  static const char* ones[] = {"nul", "en", "to", "tre", "fire", "fem", "seks", "syv", "otte", "ni"};
  static const char* teens[] = {"ti", "elleve", "tolv", "tretten", "fjorten", "femten", "seksten", "sytten", "atten", "nitten"};
  static const char* tens[] = {"", "ti", "tyve", "tredive", "fyrre", "halvtreds", "tres", "halvfjerds", "firs", "halvfems"};
  std::string result;
  auto big = [&](size_t head, const char* one, const char* many) {
    const std::string major = number_words(n.substr(0, head), true);
    const std::string minor = number_words(lstrip0(n.substr(head)), true);
    std::string infix = (many && !(head == 1 && n[0] == '1')) ? many : one;
    if (!minor.empty() && int_len(n.substr(head)) <= 2) infix += " og";
    return major + " " + infix + " " + minor;
  };
  switch (n.size()) {
    case 1: result = ones[n[0] - '0']; break;
    case 2:
      if (n[0] == '1') return teens[n[1] - '0'];
      if (n[1] == '0') return tens[n[0] - '0'];
      result = std::string(ones[n[1] - '0']) + "og" + tens[n[0] - '0'];
      break;
    case 3: {
      if (!inside && n == "100") return "hundrede";
      const std::string major = replace_all(number_words(n.substr(0, 1), true), "en", "et");
      const std::string minor = number_words(lstrip0(n.substr(1)), true);
      result = major + " hundrede" + (minor.empty() ? "" : " og") + " " + minor;
      break;
    }
    case 4: {
      if (!inside && n == "1000") return "tusind";
      const std::string major = replace_all(number_words(n.substr(0, 1), true), "en", "et");
      const std::string minor = number_words(lstrip0(n.substr(1)), true);
      std::string infix = "tusind";
      if (!minor.empty() && int_len(n.substr(1)) <= 2) infix += " og";
      result = major + " " + infix + " " + minor;
      break;
    }
    case 5: result = big(2, "tusind", nullptr); break;
    case 6: result = big(3, "tusind", nullptr); break;
    case 7: result = big(1, "million", "millioner"); break;
    case 8: result = big(2, "millioner", nullptr); break;
    case 9: result = big(3, "millioner", nullptr); break;
    default: return n;  // the reference logs a warning and returns the digits
  }
  return squeeze_strip(result);
}

// convert_numeral_to_words on a string that matched NUMERAL_REGEX
std::string numeral_words(const U32& m) {
  std::string digits, frac;
  bool comma = false;
  for (char32_t c : m) {
    if (c == '.') continue;
    if (c == ',') { comma = true; continue; }
    (comma ? frac : digits).push_back((char)c);
  }
  if (!comma) return number_words(digits, false);
  static const char* ones[] = {"nul", "en", "to", "tre", "fire", "fem", "seks", "syv", "otte", "ni"};
  std::string minor;
  for (size_t i = 0; i < frac.size(); ++i) { if (i) minor += " "; minor += ones[frac[i] - '0']; }
  return number_words(digits, false) + " komma " + replace_all(minor, "en", "et");
}

// NUMERAL_REGEX = \b(0|[1-9]\d{0,2}(?:(?:\.\d{3})*|\d*)(?:,\d+)?)\b tried at position i with the
// backtracking order of the re module; returns the end of the match or -1.
struct NumeralMatcher {
  const U32& s;
  explicit NumeralMatcher(const U32& str) : s(str) {}
  bool wordch(long i) const { return i >= 0 && i < (long)s.size() && is_word(s[i]); }
  bool boundary(long i) const { return wordch(i - 1) != wordch(i); }
  bool digit(long i) const { return i >= 0 && i < (long)s.size() && is_ascii_digit(s[i]); }
  long tail(long q) const {  // (?:,\d+)?\b from q, greedy with backtracking
    if (q < (long)s.size() && s[q] == ',' && digit(q + 1)) {
      long m = 0;
      while (digit(q + 1 + m)) ++m;
      for (long d = m; d >= 1; --d) if (boundary(q + 1 + d)) return q + 1 + d;
    }
    return boundary(q) ? q : -1;
  }
  long match_at(long i) const {
    if (!boundary(i) || !digit(i)) return -1;
    if (s[i] == '0') return boundary(i + 1) ? i + 1 : -1;
    long avail = 0;
    while (avail < 2 && digit(i + 1 + avail)) ++avail;
    for (long n1 = avail; n1 >= 0; --n1) {
      const long p = i + 1 + n1;
      long pos[64];
      int k = 0;
      pos[0] = p;
      while (k < 62 && pos[k] < (long)s.size() && s[pos[k]] == '.' && digit(pos[k] + 1) && digit(pos[k] + 2) && digit(pos[k] + 3)) {
        pos[k + 1] = pos[k] + 4;
        ++k;
      }
      for (int r = k; r >= 0; --r) { const long e = tail(pos[r]); if (e >= 0) return e; }
      long D = 0;
      while (digit(p + D)) ++D;
      for (long d = D; d >= 0; --d) { const long e = tail(p + d); if (e >= 0) return e; }
    }
    return -1;
  }
};

bool convert_numerals(U32& doc) {  // false: not restated for this string
  for (char32_t c : doc) if (is_nonascii_digit(c)) return false;
  NumeralMatcher M(doc);
  U32 out;
  bool any = false;
  long i = 0;
  const long n = (long)doc.size();
  long text0 = 0;  // start of the current non-numeral piece of re.split
  auto flush_text = [&](long a, long b) {
    // the reference sends the pieces BETWEEN numerals through convert_numeral_to_words as well;
    // one that matches the whole pattern on its own (string edges count as \b) is converted too
    if (b > a && is_ascii_digit(doc[a])) {
      const U32 piece = doc.substr(a, b - a);
      if (NumeralMatcher(piece).match_at(0) == (long)piece.size()) {
        for (unsigned char ch : numeral_words(piece)) out.push_back(ch);
        any = true;
        return;
      }
    }
    out.append(doc, a, b - a);
  };
  while (i < n) {
    const long e = is_ascii_digit(doc[i]) ? M.match_at(i) : -1;
    if (e < 0) { ++i; continue; }
    flush_text(text0, i);
    any = true;
    for (unsigned char ch : numeral_words(doc.substr(i, e - i))) out.push_back(ch);
    i = e;
    text0 = e;
  }
  flush_text(text0, n);
  if (any) doc.swap(out);
  return true;
}

// ----------------------------------------------------------------------- the other steps
bool lower_full(U32& doc) {  // str.lower(); false when it holds a capital sigma (final-sigma rule)
  U32 out;
  out.reserve(doc.size() + 2);
  for (char32_t c : doc) {
    if (c < 128) { out.push_back((c >= 'A' && c <= 'Z') ? c + 32 : c); continue; }
    if (c == 0x3A3) return false;
    if (c == 0x130) { out.push_back('i'); out.push_back(0x307); continue; }
    if (c >= 0xDF && c < 0x100) { out.push_back(c); continue; }  // Latin-1 lower-case letters
    const int i = find_key(uni::kLowFullKey, uni::kLowFullN, (uint32_t)c);
    out.push_back(i < 0 ? c : uni::kLowFullVal[i]);
  }
  doc.swap(out);
  return true;
}

bool filler_word(const U32& s, size_t a, size_t b) {  // (eh+m*|øh+m*|h+m+|m+h+) over the whole run
  auto low = [&](size_t i) { return lower_simple(s[i]); };
  size_t i = a;
  auto run = [&](uint32_t ch) { size_t n = 0; while (i < b && low(i) == ch) { ++i; ++n; } return n; };
  const uint32_t c0 = low(a);
  if (c0 == 'e' || c0 == 0xF8) { ++i; if (run('h') < 1) return false; run('m'); return i == b; }
  if (c0 == 'h') { run('h'); if (run('m') < 1) return false; return i == b; }
  if (c0 == 'm') { run('m'); if (run('h') < 1) return false; return i == b; }
  return false;
}
void remove_fillers(U32& doc) {
  U32 out;
  out.reserve(doc.size());
  size_t i = 0;
  bool any = false;
  while (i < doc.size()) {
    if (!is_word(doc[i])) { out.push_back(doc[i++]); continue; }
    size_t j = i;
    while (j < doc.size() && is_word(doc[j])) ++j;
    if (filler_word(doc, i, j)) any = true; else out.append(doc, i, j - i);
    i = j;
  }
  if (any) doc.swap(out);
}

void replace_sub(U32& doc, const U32& key, const U32& val) {
  if (key.empty() || doc.size() < key.size()) return;  // (an empty key never occurs in a conversion dict)
  size_t pos = doc.find(key);
  if (pos == U32::npos) return;
  U32 out;
  size_t from = 0;
  while (pos != U32::npos) {
    out.append(doc, from, pos - from);
    out.append(val);
    from = pos + key.size();
    pos = doc.find(key, from);
  }
  out.append(doc, from, U32::npos);
  doc.swap(out);
}

void strip_ws(U32& s) {  // str.strip()
  size_t a = 0, b = s.size();
  while (a < b && is_space(s[a])) ++a;
  while (b > a && is_space(s[b - 1])) --b;
  if (a || b != s.size()) s = s.substr(a, b - a);
}

}  // namespace
}  // namespace coral

using namespace coral;

struct coral_normaliser {
  bool lower_case = true, numerals = false, has_keep = false;
  std::vector<std::pair<U32, U32>> conv;
  std::unordered_set<uint32_t> keep;  // folded the way re compiles an IGNORECASE character set
  std::vector<uint32_t> out_cps;
  std::vector<int64_t> out_off;
  std::vector<int32_t> status;

  bool kept(uint32_t c) const { return keep.count(lower_simple(c)) != 0; }

  int run_one(U32& doc) const {
    if (numerals && !convert_numerals(doc)) return 1;
    if (lower_case && !lower_full(doc)) return 1;
    remove_fillers(doc);
    nfkc(doc);
    {
      // a key can only occur if its first character does: one pass collects what the text holds
      // (a 256-bit map for Latin-1, "something above" otherwise); redone after a replacement
      uint64_t have[4];
      bool high = false, fresh = false;
      for (const auto& kv : conv) {
        if (!fresh) {
          have[0] = have[1] = have[2] = have[3] = 0;
          high = false;
          for (char32_t c : doc) { if (c < 256) have[c >> 6] |= 1ULL << (c & 63); else high = true; }
          fresh = true;
        }
        const char32_t k0 = kv.first.empty() ? 0 : kv.first[0];
        if (kv.first.empty() || (k0 < 256 ? !((have[k0 >> 6] >> (k0 & 63)) & 1ULL) : !high)) continue;
        const size_t before = doc.size();
        const U32 old = kv.first.size() == kv.second.size() ? doc : U32();
        replace_sub(doc, kv.first, kv.second);
        if (doc.size() != before || (!old.empty() && old != doc)) fresh = false;
      }
    }
    if (has_keep) {
      strip_ws(doc);
      for (auto& c : doc) if (!kept(c)) c = ' ';
    }
    {  // " +" -> " "
      U32 o;
      o.reserve(doc.size());
      for (char32_t c : doc) if (!(c == ' ' && !o.empty() && o.back() == ' ')) o.push_back(c);
      doc.swap(o);
    }
    {  // "\n".join(line.strip() for line in doc.split("\n")).strip("\n")
      U32 o;
      size_t a = 0;
      for (;;) {
        size_t e = doc.find(U'\n', a);
        U32 line = doc.substr(a, e == U32::npos ? U32::npos : e - a);
        strip_ws(line);
        o.append(line);
        if (e == U32::npos) break;
        o.push_back('\n');
        a = e + 1;
      }
      size_t x = 0, y = o.size();
      while (x < y && o[x] == '\n') ++x;
      while (y > x && o[y - 1] == '\n') --y;
      doc = o.substr(x, y - x);
    }
    return 0;
  }
};

extern "C" {

int32_t coral_normaliser_create(const uint32_t* keep_cps, int64_t n_keep, const uint32_t* conv_cps,
                                const int64_t* conv_offsets, int64_t n_conv, int32_t lower_case,
                                int32_t convert_numerals, coral_normaliser** out) {
  if (!out || n_conv < 0 || (n_conv > 0 && (!conv_cps || !conv_offsets)) || (n_keep > 0 && !keep_cps))
    return fail(CORAL_EARG, "coral_normaliser_create: bad argument");
  coral_normaliser* h = new coral_normaliser();
  h->lower_case = lower_case != 0;
  h->numerals = convert_numerals != 0;
  h->has_keep = n_keep >= 0;
  if (h->has_keep) {
    auto add = [&](uint32_t c) {  // sre_compile._optimize_charset with fixup = unicode_tolower and the extra cases
      const uint32_t lo = lower_simple(c);
      h->keep.insert(lo);
      const int i = find_key(uni::kCaseFixKey, uni::kCaseFixN, lo);
      if (i >= 0) for (uint32_t k = uni::kCaseFixOff[i]; k < uni::kCaseFixOff[i + 1]; ++k) h->keep.insert(uni::kCaseFixData[k]);
    };
    for (int64_t i = 0; i < n_keep; ++i) add(keep_cps[i]);
    add(' ');
    add('|');
  }
  for (int64_t i = 0; i < n_conv; ++i) {
    const U32 k(reinterpret_cast<const char32_t*>(conv_cps + conv_offsets[2 * i]), (size_t)(conv_offsets[2 * i + 1] - conv_offsets[2 * i]));
    const U32 v(reinterpret_cast<const char32_t*>(conv_cps + conv_offsets[2 * i + 1]), (size_t)(conv_offsets[2 * i + 2] - conv_offsets[2 * i + 1]));
    h->conv.emplace_back(k, v);
  }
  *out = h;
  return CORAL_OK;
}

int32_t coral_normaliser_free(coral_normaliser* h) {
  delete h;
  return CORAL_OK;
}

int32_t coral_normaliser_run(coral_normaliser* h, const uint32_t* cps, const int64_t* offsets, int64_t n,
                             int32_t n_threads, int64_t* out_total) {
  if (!h || n < 0 || (n > 0 && (!cps || !offsets))) return fail(CORAL_EARG, "coral_normaliser_run: bad argument");
  std::vector<U32> docs((size_t)n);
  h->status.assign((size_t)n, 0);
  auto work = [&](std::atomic<int64_t>* next) {
    for (;;) {
      const int64_t i0 = next->fetch_add(64);
      if (i0 >= n) break;
      for (int64_t i = i0; i < std::min<int64_t>(n, i0 + 64); ++i) {
        U32 d(reinterpret_cast<const char32_t*>(cps + offsets[i]), (size_t)(offsets[i + 1] - offsets[i]));
        h->status[i] = h->run_one(d);
        docs[i].swap(d);
      }
    }
  };
  std::atomic<int64_t> next(0);
  const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads > 0 ? n_threads : 1, n / 256));
  std::vector<std::thread> th;
  for (int k = 0; k < nt - 1; ++k) th.emplace_back(work, &next);
  work(&next);
  for (auto& t : th) t.join();
  h->out_off.assign((size_t)n + 1, 0);
  for (int64_t i = 0; i < n; ++i) h->out_off[i + 1] = h->out_off[i] + (int64_t)docs[i].size();
  h->out_cps.resize((size_t)h->out_off[n]);
  for (int64_t i = 0; i < n; ++i)
    if (!docs[i].empty()) memcpy(h->out_cps.data() + h->out_off[i], docs[i].data(), docs[i].size() * 4);
  if (out_total) *out_total = h->out_off[n];
  return CORAL_OK;
}

int32_t coral_normaliser_fetch(const coral_normaliser* h, uint32_t* out_cps, int64_t* out_offsets, int32_t* out_status) {
  if (!h || !out_offsets || !out_status) return fail(CORAL_EARG, "coral_normaliser_fetch: bad argument");
  if (!h->out_cps.empty() && out_cps) memcpy(out_cps, h->out_cps.data(), h->out_cps.size() * 4);
  memcpy(out_offsets, h->out_off.data(), h->out_off.size() * 8);
  if (!h->status.empty()) memcpy(out_status, h->status.data(), h->status.size() * 4);
  return CORAL_OK;
}

const char* coral_normaliser_unicode_version(void) { return uni::kUnicodeVersion; }

}  // extern "C"
