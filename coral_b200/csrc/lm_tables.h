// GPU-resident n-gram LM and lexicon: table layouts and the query routines.
//
// Replaces, for the decode hot path, kenlm.Model.BaseScore / BeginSentenceWrite /
// NullContextWrite / __contains__ (UP:kenlm lm/model.cc, python/kenlm.pyx -- not on
// disk; semantics as specified in SURVEY.md section 8 A8) and pyctcdecode's unigram set /
// pygtrie.CharTrie prefix test (UP:pyctcdecode language_model.py).
// Reference call sites: R:src/coral/ngram.py:341-343.
//
// Layout in HBM (one copy per GPU; see DESIGN.md "LM tables"):
//   uni[wid]      8 B   {float prob, float backoff}        direct-indexed by word id (0 = <unk>)
//   ng[slot]     16 B   {u64 chain key, float prob, float backoff}   ONE open-addressing
//                       table for all orders >= 2, linear probing, load <= 0.5, key 0 = empty.
//                       chain key of "c2 c1 w" = combine(combine(id(w), c1), c2): KenLM's probing
//                       keys, built in the order a longest-match lookup extends them (one probe
//                       per context word, stop at the first miss).
//   lex[slot]    16 B   {u64 word hash, u32 word id, u32 flags}      every prefix of every
//                       word of (LM vocabulary U unigram list); rolling hash over code points.
//
// Everything here compiles as plain C++ too (tests/hostsim builds the same routines for
// the CPU to check the algorithm without a GPU; the product never runs that build).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CORAL_HD __host__ __device__ __forceinline__
#else
#define CORAL_HD inline
#endif

namespace coral {

constexpr int kMaxOrder = 6;
constexpr int kMaxCtx = kMaxOrder - 1;

// lexicon flags
constexpr uint32_t kLexPrefixOfUnigram = 1u;  // a prefix (proper or not) of a unigram-set word
constexpr uint32_t kLexInUnigrams = 2u;       // the whole string is in the unigram set
constexpr uint32_t kLexInLm = 4u;             // the whole string is in the LM vocabulary (id != 0)

struct UniEntry {
  float prob;
  float backoff;
};
struct alignas(16) NgSlot {
  uint64_t key;
  float prob;
  float backoff;
};
struct alignas(16) LexSlot {
  uint64_t key;
  uint32_t wid;
  uint32_t flags;
};

struct LmState {
  uint32_t w[kMaxCtx];
  float b[kMaxCtx];
  uint32_t len;
};

// Read-only view handed to kernels (device pointers) or to the host simulation (host pointers).
struct LmView {
  const UniEntry* uni;
  const NgSlot* ng;
  const LexSlot* lex;
  uint64_t ng_mask;   // table size - 1 (power of two); 0 when order == 1
  uint64_t lex_mask;  // table size - 1
  uint32_t n_vocab;
  int32_t order;
  uint32_t bos_id;
  uint32_t eos_id;
  int32_t has_unigrams;  // len(unigram_set) > 0 after filtering with the LM vocabulary
  int32_t kenlm_keys;    // informational: the tables came from a KenLM binary (keys are KenLM's either way)
  int32_t present;       // 0 => decoder built without a language model
  float score_ub;        // upper bound of any log10 probability the model returns (>= 0)
  // heavy-frame path: per lexicon slot, the alphabet tokens whose label extends that prefix to
  // another prefix of a unigram-set word (no partial-word penalty); NULL = not built
  const uint64_t* lex_ok;
  uint64_t root_ok;      // the same for the empty prefix
};

CORAL_HD uint64_t mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}

constexpr uint64_t kWordHashSeed = 0x243F6A8885A308D3ULL;

// rolling hash of a word, one Unicode code point at a time
// Incremental hashes sit on dependent chains of the beam kernel's critical path, so each
// step is one 64-bit multiply and one xor-shift (not a full finaliser); the builder checks
// that no two distinct words / n-grams collide.
CORAL_HD uint64_t hash_step(uint64_t h, uint64_t x) {
  h = (h ^ (x + 0x9E3779B97F4A7C15ULL)) * 0xff51afd7ed558ccdULL;
  h ^= h >> 32;
  return h ? h : 1;  // 0 is the empty-slot marker
}
CORAL_HD uint64_t word_hash_push(uint64_t h, uint32_t cp) { return hash_step(h, (uint64_t)cp); }
// n-gram identity = KenLM's own chain hash (UP:kenlm lm/search_hashed.hh, CombineWordHash): the
// predicted word's id, then one multiply-xor per context word, most recent first. A model read
// from a KenLM probing binary only has these keys (the words of an n-gram are not stored), so
// ARPA-built tables use the same scheme and the kernel has one way to form a key.
CORAL_HD uint64_t kenlm_combine(uint64_t k, uint32_t w) {
  return (k * 8978948897894561157ULL) ^ ((uint64_t)(1u + w) * 17894857484156487943ULL);
}
// one key scheme for every model (no per-probe branch in the kernel): KenLM's
CORAL_HD uint64_t ng_key_first(const LmView&, uint32_t w) { return (uint64_t)w; }
CORAL_HD uint64_t ng_key_next(const LmView&, uint64_t k, uint32_t w) { return kenlm_combine(k, w); }

CORAL_HD bool lex_find(const LmView& lm, uint64_t h, uint32_t& wid, uint32_t& flags) {
  uint64_t i = (h >> 20) & lm.lex_mask;
  for (;;) {
#if defined(__CUDA_ARCH__)
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(lm.lex + i));
    const uint64_t key = ((uint64_t)v.y << 32) | v.x;
    if (key == h) { wid = v.z; flags = v.w; return true; }
#else
    const LexSlot s = lm.lex[i];
    const uint64_t key = s.key;
    if (key == h) { wid = s.wid; flags = s.flags; return true; }
#endif
    if (key == 0) return false;
    i = (i + 1) & lm.lex_mask;
  }
}

// slot index of a prefix in the lexicon table, -1 if absent (heavy-frame path only)
CORAL_HD long long lex_slot(const LmView& lm, uint64_t h) {
  uint64_t i = (h >> 20) & lm.lex_mask;
  for (;;) {
    const uint64_t key = lm.lex[i].key;
    if (key == h) return (long long)i;
    if (key == 0) return -1;
    i = (i + 1) & lm.lex_mask;
  }
}

CORAL_HD bool ng_find(const LmView& lm, uint64_t k, float& prob, float& backoff) {
  uint64_t i = (k >> 20) & lm.ng_mask;
  for (;;) {
#if defined(__CUDA_ARCH__)
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(lm.ng + i));
    const uint64_t key = ((uint64_t)v.y << 32) | v.x;
    if (key == k) { prob = __uint_as_float(v.z); backoff = __uint_as_float(v.w); return true; }
#else
    const NgSlot s = lm.ng[i];
    const uint64_t key = s.key;
    if (key == k) { prob = s.prob; backoff = s.backoff; return true; }
#endif
    if (key == 0) return false;
    i = (i + 1) & lm.ng_mask;
  }
}

CORAL_HD float f32_add(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b;
  return r;
#endif
}

// kenlm BaseScore: log10 p(w | in) in float32 and the out state.
// Longest match found by extending the chain key one context word at a time and
// stopping at the first absent n-gram; then the unused context back-offs are added
// in float32 in ascending context length.
// The chain keys depend only on the words, so the device path first prefetches every order's
// slot into L1 and then walks the dependent chain against L1. Loops are kept rolled: this routine sits
// on the per-frame path of the beam kernel, whose code must stay instruction-cache sized.
CORAL_HD float lm_base_score(const LmView& lm, const LmState& in, uint32_t w, LmState& out,
                             int* probes = nullptr) {
  const uint32_t nctx = in.len < (uint32_t)(lm.order - 1) ? in.len : (uint32_t)(lm.order > 0 ? lm.order - 1 : 0);
#if defined(__CUDA_ARCH__)
  {
    // first pass: every order's slot is prefetched into L1 (the keys depend only on the words)
    uint64_t kp = ng_key_first(lm, w);
#pragma unroll 1
    for (uint32_t i = 0; i < nctx; ++i) {
      kp = ng_key_next(lm, kp, in.w[i]);
      asm volatile("prefetch.global.L1 [%0];" ::"l"(lm.ng + ((kp >> 20) & lm.ng_mask)));
    }
  }
#endif
  const UniEntry u = lm.uni[w];
  float prob = u.prob;
  out.w[0] = w;
  out.b[0] = u.backoff;
  uint32_t olen = 1;
  uint32_t ngram_len = 1;
  uint64_t key = ng_key_first(lm, w);
  int np = 1;
#pragma unroll 1
  for (uint32_t i = 0; i < nctx; ++i) {
    const int n = (int)i + 2;
    key = ng_key_next(lm, key, in.w[i]);
    float p, b;
    ++np;
    if (!ng_find(lm, key, p, b)) break;
    prob = p;
    ngram_len = (uint32_t)n;
    if (n < lm.order) {
      out.w[olen] = in.w[i];
      out.b[olen] = b;
      ++olen;
    }
  }
#pragma unroll 1
  for (uint32_t i = ngram_len - 1; i < in.len; ++i) prob = f32_add(prob, in.b[i]);
  const uint32_t keep = (uint32_t)(lm.order - 1);
  out.len = olen < keep ? olen : keep;
  if (probes) *probes = np;
  return prob;
}

CORAL_HD void lm_begin_sentence(const LmView& lm, LmState& s) {
  s.len = 1;
  s.w[0] = lm.bos_id;
  s.b[0] = lm.uni[lm.bos_id].backoff;
  if (lm.order < 2) s.len = 0;
}

CORAL_HD void lm_null_context(LmState& s) { s.len = 0; }

}  // namespace coral
