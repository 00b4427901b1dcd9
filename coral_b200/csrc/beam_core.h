// Prefix beam search with word-level n-gram shallow fusion: one utterance per thread group.
//
// Replaces pyctcdecode 0.5.0 BeamSearchDecoderCTC._decode_logits / _merge_beams /
// _get_lm_beams / _sort_and_trim_beams and LanguageModel.score / score_partial_token
// (UP: pyctcdecode decoder.py, language_model.py -- not on disk; behaviour as specified in
// SURVEY.md section 8 A5/A7). Reference call sites: R:src/coral/ngram.py:341-343,
// HF:models/wav2vec2_with_lm/processing_wav2vec2_with_lm.py:398-406, :565-572,
// HF:pipelines/automatic_speech_recognition.py:612-616.
//
// How the reference's Python objects map onto the device:
//   * a text prefix is a node of a per-utterance character trie in HBM (parent, token);
//     spaces that do not close a word never create nodes, so "equal text" <=> "equal node";
//   * a beam is (node, last_char, logit_score) + cached per-node word state (rolling word
//     hash, lexicon flags, LM word id, LM boundary record, raw LM score);
//   * pyctcdecode's dict merge keyed on (text, word_part, last_char) becomes a gather: for
//     every live node m and kept token c the (at most four) parent beams that collapse onto
//     (m, c) are found through one small shared-memory hash of the live nodes, so no
//     floating-point atomics are needed and the log-sum-exp runs in the reference's order;
//   * the text-keyed LM cache becomes "a word-boundary node keeps its LM record";
//   * heapq.nlargest (stable) becomes rank-by-(score desc, first-candidate-index asc).
//
// The code is phase-structured: inside CORAL_LANES(...) lanes touch only their own items
// and communicate through shared memory + atomics; GSYNC separates phases. Compiled by
// nvcc it is the kernel body; compiled with -DCORAL_HOSTSIM (tests/hostsim only) the lane
// loop runs sequentially on the CPU so the algorithm can be checked without a GPU.
#pragma once
#include <stddef.h>
#include <math.h>
#include <stdint.h>

#include "lm_tables.h"

namespace coral {

#if defined(CORAL_HOSTSIM)
#define CORAL_DEV inline
#define CORAL_DEV_OUTLINE inline
template <class T>
inline T atom_add(T* p, T v) { T o = *p; *p = o + v; return o; }
inline unsigned long long atom_max_u64(unsigned long long* p, unsigned long long v) {
  unsigned long long o = *p; if (v > o) *p = v; return o;
}
inline unsigned long long atom_cas_u64(unsigned long long* p, unsigned long long c, unsigned long long v) {
  unsigned long long o = *p; if (o == c) *p = v; return o;
}
inline uint32_t atom_cas_u32(uint32_t* p, uint32_t c, uint32_t v) { uint32_t o = *p; if (o == c) *p = v; return o; }
inline uint32_t atom_exch_u32(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = v; return o; }
inline uint32_t atom_add_u16(uint16_t* p, uint32_t v) { uint32_t o = *p; *p = (uint16_t)(o + v); return o; }
inline void atom_or_u64(unsigned long long* p, unsigned long long v) { *p |= v; }
#define CORAL_LANES(NT) for (int lane = 0; lane < (NT); ++lane)
#define CORAL_GSYNC(NT) ((void)0)
#define CORAL_WSYNC() ((void)0)
#else
#define CORAL_DEV __device__ __forceinline__
// bulky or rarely executed paths are kept out of line so that the per-frame loop stays
// inside the instruction cache (the fully inlined kernel was 157 KB of SASS)
#define CORAL_DEV_OUTLINE __device__ __noinline__
template <class T>
__device__ __forceinline__ T atom_add(T* p, T v) { return atomicAdd(p, v); }
__device__ __forceinline__ unsigned long long atom_max_u64(unsigned long long* p, unsigned long long v) {
  return atomicMax(p, v);
}
__device__ __forceinline__ unsigned long long atom_cas_u64(unsigned long long* p, unsigned long long c,
                                                           unsigned long long v) {
  return atomicCAS(p, c, v);
}
__device__ __forceinline__ uint32_t atom_cas_u32(uint32_t* p, uint32_t c, uint32_t v) { return atomicCAS(p, c, v); }
__device__ __forceinline__ uint32_t atom_exch_u32(uint32_t* p, uint32_t v) { return atomicExch(p, v); }
__device__ __forceinline__ void atom_or_u64(unsigned long long* p, unsigned long long v) { atomicOr(p, v); }
// 16-bit counters packed two per word: add to the right half, return that half's old value
__device__ __forceinline__ uint32_t atom_add_u16(uint16_t* p, uint32_t v) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  unsigned int* w = reinterpret_cast<unsigned int*>(a & ~(uintptr_t)3);
  const unsigned sh = (a & 2) ? 16u : 0u;
  const unsigned old = atomicAdd(w, v << sh);
  return (old >> sh) & 0xFFFFu;
}
template <int NT>
__device__ __forceinline__ void group_sync() {
  if (NT == 32) {
    __syncwarp();
  } else {
    __syncthreads();  // one thread group per CTA
  }
}
#define CORAL_LANES(NT) for (int lane = (int)(threadIdx.x % (NT)), _once = 1; _once; _once = 0)
#define CORAL_GSYNC(NT) group_sync<NT>()
#define CORAL_WSYNC() __syncwarp()
#endif

// ---------------------------------------------------------------- fp64 with fixed rounding
CORAL_HD double d_add(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
CORAL_HD double d_mul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
CORAL_HD double d_div(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __ddiv_rn(a, b);
#else
  return a / b;
#endif
}
CORAL_HD uint32_t mulhi_u32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((unsigned long long)a * b) >> 32);
#endif
}
// pyctcdecode _sum_log_scores: math.log / math.exp in float64
CORAL_HD double sum_log_scores(double s1, double s2) {
  if (s1 >= s2) return d_add(s1, log(d_add(1.0, exp(d_add(s2, -s1)))));
  return d_add(s2, log(d_add(1.0, exp(d_add(s1, -s2)))));
}
CORAL_HD unsigned long long ordered_u64(double x) {
  union { double d; unsigned long long u; } c;
  c.d = x;
  return (c.u >> 63) ? ~c.u : (c.u | 0x8000000000000000ULL);
}

// ---------------------------------------------------------------------------- parameters
constexpr int kVMax = 64;        // alphabet size supported by this build
#ifndef CORAL_KCHUNK
#define CORAL_KCHUNK 12
#endif
constexpr int kChunk = CORAL_KCHUNK;  // frames staged in shared memory per pass
constexpr int kMaxLabelCps = 8;  // code points per alphabet label
constexpr uint32_t kNone16 = 0xFFFFu;
constexpr uint32_t kNoTok = 0xFFu;
constexpr uint32_t kNoNode = 0xFFFFFFFFu;
constexpr uint32_t kListEnd = 0xFFFFu;

// beam flags (4 bits in meta)
constexpr uint32_t kOovPartial = 1u;  // word_part is not a prefix of any unigram-set word
constexpr uint32_t kDead = 2u;        // word_part is not a prefix of anything in the lexicon
constexpr uint32_t kInUni = 4u;       // word_part is a unigram-set word
constexpr uint32_t kInLm = 8u;        // word_part is in the LM vocabulary

struct DecodeParams {
  int32_t V;
  int32_t blank_id;
  int32_t space_id;
  int32_t beam_width;
  int32_t n_best;       // beams written per utterance (<= beam_width)
  int32_t T_max;        // row pitch of logits [B, T_max, V] and of out_tokens
  int32_t input_mode;   // 0 auto (pyctcdecode's test, evaluated per utterance), 1 logits, 2 probabilities
  int32_t host_input;   // logits live in pinned HOST memory: 1 = uncached loads, 2 = plain loads + L2 prefetch
  int32_t prune_history;  // pyctcdecode prune_history=True: n = max(1, LM order - 1) words of history, 0 = off
  int32_t score_boundary;
  float token_min_logp;  // compared in float32
  double beam_prune_logp;
  double alpha, beta, unk_score_offset;
  double log_base_change;  // 1 / log10(e) as the reference computes it
  double bucket_scale;     // score buckets per unit of score (filled by set_bucket_scale on the host)
  uint32_t label_cps[kVMax][kMaxLabelCps];
  uint8_t label_ncp[kVMax];
};

struct BndRec {  // LM record of a word boundary (a complete-words text prefix)
  double lm_raw;
  LmState st;
};

// prune_history: the last words of a complete-words text prefix (64-bit hashes of the words, most
// recent first, 0 = none) and the hash of the n most recent ones. One record per LM boundary
// record (same index), kept in HBM and only when prune_history is on.
struct HistRec {
  unsigned long long H;
  unsigned long long w[kMaxCtx];
};

// Candidates ("outputs") of one frame, struct-of-arrays so that ranking streams over a dense
// array of 64-bit order-preserving score keys.
struct OutView {
  unsigned long long* key;  // ordered_u64(combined score)
  double* logit;
  uint32_t* order;  // first candidate index in (token-major, beam-minor) order
  uint32_t* aux;    // kind 2: LM word id ; kind 3: boundary record
  uint32_t* child;  // kind 1: representative beam of the live child ; kind 3: node id ; final: text node
  uint32_t* info;   // source beam (16 bits) | token << 16 | kind/flags << 24
};

// One finished word of a beam's text: frames [start, end) and the record of the word before it.
// pyctcdecode carries text_frames as a Python list per beam; beams share list prefixes, so
// an append-only arena of back-linked records holds them all (index 0 = empty list).
struct alignas(16) FrameRec {
  uint32_t parent;
  int32_t start, end;
  uint32_t pad;
};

struct SlotScratch {  // per thread-group arenas in HBM, reused utterance after utterance
  FrameRec* wf;         // word-frame records (only with word frames enabled)
  uint32_t wf_cap;
  uint32_t* node_parent;
  uint32_t* node_info;  // tok | bnd << 8
  BndRec* bnd;
  OutView outs_g;       // overflow for frames with more candidates than fit in smem
  float* rowsum;        // [T_max] row sums of the utterance being classified
  // rarely used arenas sit right behind the overflow candidates (outs_g.info) and are addressed
  // from there, so that they cost the hot path no registers:
  //   hv_sorted u32 [outs_cap]  heavy frames: (upper-bound bin << 24 | item id), best bins first
  //   hv_masks  u64 [1024]      heavy frames: per live-prefix-hash slot, the tokens whose extension is live
  //   hist      HistRec [bnd_cap] with prune_history
  static constexpr size_t kHvMaskBytes = 1024 * 8;
  CORAL_HD uint32_t* hv_sorted() const { return outs_g.info + outs_cap; }
  CORAL_HD unsigned long long* hv_masks() const {
    return reinterpret_cast<unsigned long long*>(outs_g.info + 2 * (size_t)outs_cap);
  }
  //   hv_bin    u8  [outs_cap]  heavy frames: upper-bound bin of every item (0xFF = no candidate)
  CORAL_HD uint8_t* hv_bin() const { return reinterpret_cast<uint8_t*>(hv_masks()) + kHvMaskBytes; }
  CORAL_HD HistRec* hist() const {
    return reinterpret_cast<HistRec*>(hv_bin() + (((size_t)outs_cap + 15) & ~(size_t)15));
  }
  uint32_t node_cap, bnd_cap, outs_cap;
};

struct UttIO {
  const float* logits;  // [T, V] of this utterance
  int32_t T;
  // outputs
  int32_t* out_n;       // scalar: number of final beams
  double* out_logit;    // [n_best]
  double* out_comb;     // [n_best]
  uint8_t* out_tokens;  // [n_best, T_max]
  int32_t* out_len;     // [n_best]
  int32_t* out_frames;  // [n_best, max_words, 2] word frames (start, end) -- word-frame kernels only
  int32_t* out_nwords;  // [n_best]
  int32_t max_words;
  int32_t* out_status;  // scalar: 0 ok, -4 capacity
  // optional [16]: extensions, LM scorings, n-gram probes, frames, lexicon probes, back-pointer records,
  // LM boundary records, (unused); [8..15] cycles per phase (thread 0): hash, expand,
  // overflow select, bucket, scatter, rank+commit, grow, frame staging
  unsigned long long* stats;
};

#ifndef CORAL_KNB
#define CORAL_KNB 128
#endif
constexpr int kNB = CORAL_KNB;      // score buckets over the prune window (<= 256, multiple of 8)
// the buckets span the prune window plus a margin; any monotone map keeps the ranking exact
inline void set_bucket_scale(DecodeParams& P) {
  P.bucket_scale = (double)kNB / ((P.beam_prune_logp < 0.0 ? -P.beam_prune_logp : 0.0) + 4.0);
}

// per-beam word timing (pyctcdecode's part_frames and text_frames), present only in the
// word-frame instantiation of the kernel
template <int BW, bool FRAMES>
struct WordFrames {
  int32_t pf0[2][BW], pf1[2][BW];  // part_frames (start, end) of the open word, -1 = unset
  uint32_t head[2][BW];            // last FrameRec of text_frames
  uint32_t count;                  // records used in SlotScratch::wf
};
template <int BW>
struct WordFrames<BW, false> {};

// cycle timers of selected device operations: only in the instrumented (STATS) instantiation
template <bool STATS>
struct OpTimers {
  unsigned long long opc[8];  // tuning: cycles spent inside selected device operations
  uint32_t opn[8];            //         and how many times each ran
};
template <>
struct OpTimers<false> {};

// heavy frames (expand_heavy): histogram of the items' score upper bounds, two buckets per bin;
// only in the instantiations that contain the heavy-frame path
template <bool HEAVY>
struct HeavyArea {
  uint16_t ubcnt[kNB / 2];
};
template <>
struct HeavyArea<false> {};

template <int BW, int OUTC, bool FRAMES = false, bool STATS = true, bool HEAVY = true>
struct GroupShared {
  WordFrames<BW, FRAMES> wf;
  OpTimers<STATS> tm;
  static constexpr int HS = BW <= 32 ? 64 : (BW <= 64 ? 128 : (BW <= 128 ? 256 : (BW <= 256 ? 512 : 1024)));  // >= 2 BW
  // beams, double buffered
  double logit[2][BW];
  double lm_raw[2][BW];
  unsigned long long whash[2][BW];
  unsigned long long nh[2][BW];  // identity of the text prefix: 64-bit hash of its token string
  unsigned long long ph[2][BW];  // identity of its parent prefix
  uint32_t node[2][BW];          // back-pointer arena index (text output only)
  uint32_t bnd[2][BW];
  uint32_t wid[2][BW];
  uint32_t meta[2][BW];  // tok | lc << 8 | flags << 16
  uint16_t wlen[2][BW];
  // live-node hash
  unsigned long long hkey[HS];
  uint16_t sb0[HS], sb1[HS];
  uint16_t ne_slot[BW];
  // candidates of this frame (also reused as the 256-bin histogram of the overflow path)
  unsigned long long o_key[OUTC];
  double o_logit[OUTC];
  uint32_t o_order[OUTC];
  uint32_t o_aux[OUTC];
  uint32_t o_child[OUTC];
  uint32_t o_info[OUTC];
  // score buckets for ranking (phase 3): per bucket a count and a linked list of its candidates
  // (built while the candidates are produced), plus the exclusive prefix sums of the counts
  uint32_t bcnt[kNB];
  uint32_t bhead[kNB];      // first candidate of the bucket's list (kListEnd = empty)
  uint16_t bstart[kNB];     // candidates in strictly better buckets (capped at 0xFFFF)
  uint16_t o_next[OUTC];    // next candidate of the same bucket
  uint8_t o_bkt[OUTC];      // its bucket
  // staged frames
  float lp[kChunk][kVMax];
  uint8_t kept[kChunk][kVMax];
  uint8_t nkept[kChunk];
  uint8_t amax[kChunk];
  // scalars; the per-frame counters are double-buffered by frame parity so that the next
  // frame's copy can be cleared without an extra barrier
  unsigned long long gmax[2];
  double mhat;  // reference score of this frame's buckets: an estimate of its best score
  unsigned long long sel_prefix, sel_mask;
  uint32_t nN[2], n_out[2], S[2];
  uint32_t node_count, bnd_count;
  uint32_t sel_need, sel_eq, sel_cut, sel_n;
  uint32_t gsum[16];
  int32_t status, utt, is_prob;
  uint32_t cnt[8];  // work counters of this utterance (flushed to UttIO::stats at its end)
  HeavyArea<HEAVY> hv;
};

CORAL_HD uint32_t meta_pack(uint32_t tok, uint32_t lc, uint32_t flags) { return tok | (lc << 8) | (flags << 16); }
CORAL_HD uint32_t meta_tok(uint32_t m) { return m & 0xFFu; }
CORAL_HD uint32_t meta_lc(uint32_t m) { return (m >> 8) & 0xFFu; }
CORAL_HD uint32_t meta_flags(uint32_t m) { return (m >> 16) & 0xFFu; }
constexpr uint32_t kLcNone = 0xFFu;   // last_char None (start of utterance)
constexpr uint32_t kLcBlank = 0xFEu;  // last_char "" (blank)

// Identity of a text prefix = 64-bit hash of its (space-normalised) token string, rolled one
// token at a time. Two prefixes are "the same text" iff their hashes are equal -- the same
// kind of guarantee KenLM gives for n-gram identity (64-bit hashes, no stored strings); with
// ~1e4 prefixes per utterance the collision probability is ~3e-12 per utterance. This
// replaces a per-utterance (parent, token) -> node table in HBM whose dependent
// compare-and-swap round trips (~6k cycles each) dominated the frame time.
constexpr unsigned long long kRootHash = 0x452821E638D01377ULL;
CORAL_HD unsigned long long child_hash(unsigned long long h, uint32_t tok) {
  // two multiply/xor-shift rounds: this hash IS the identity, so it gets the better mixing
  h = (h ^ (0x9E3779B97F4A7C15ULL * (unsigned long long)(tok + 1u))) * 0xff51afd7ed558ccdULL;
  h ^= h >> 33;
  h *= 0xc4ceb9fe1a85ec53ULL;
  h ^= h >> 29;
  return h ? h : 1ULL;  // 0 marks an empty hash slot
}

// pyctcdecode LanguageModel.score_partial_token, hotwords empty
static CORAL_DEV_OUTLINE double partial_score_long(double u, uint32_t wlen) { return d_div(d_mul(u, (double)wlen), 6.0); }
CORAL_DEV double partial_score(const DecodeParams& P, uint32_t wlen, uint32_t flags) {
  if (wlen == 0) return 0.0;
  const double u = d_mul(P.unk_score_offset, (flags & kOovPartial) ? 1.0 : 0.0);
  return wlen > 6 ? partial_score_long(u, wlen) : u;
}

// pyctcdecode LanguageModel.score: alpha * log10-score * ln10 + beta
static CORAL_DEV_OUTLINE float lm_base_score_call(const LmView& lm, const LmState& in, uint32_t w, LmState& out,
                                                  int* probes) {
  return lm_base_score(lm, in, w, out, probes);
}
static CORAL_DEV_OUTLINE double lm_word_score(const LmView& lm, const DecodeParams& P, const LmState& in, uint32_t wid,
                               bool oov, bool is_last, LmState& out, uint32_t* cnt) {
  int np = 0;
  double x = (double)lm_base_score_call(lm, in, wid, out, &np);
  if (oov) x = d_add(x, P.unk_score_offset);
  if (is_last) {
    double e = 0.0;
    if (P.score_boundary) {
      LmState tmp;
      int np2 = 0;
      e = (double)lm_base_score_call(lm, out, lm.eos_id, tmp, &np2);
      np += np2;
    }
    x = d_add(x, e);
  }
  if (cnt) { atom_add(&cnt[1], 1u); atom_add(&cnt[2], (uint32_t)np); }
  return d_add(d_mul(d_mul(P.alpha, x), P.log_base_change), P.beta);
}

// numpy's float32 pairwise summation (np.add.reduce over a contiguous axis): n < 8 sequential;
// n <= 128 eight strided accumulators combined as a balanced tree, remainder in order; larger n
// split in halves (the left half rounded down to a multiple of 8), recursively.
CORAL_HD float np_pairwise_leaf(const float* a, int n) {
  if (n < 8) {
    float r = 0.0f;
    for (int i = 0; i < n; ++i) r = f32_add(r, a[i]);
    return r;
  }
  float r[8];
  for (int j = 0; j < 8; ++j) r[j] = a[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int j = 0; j < 8; ++j) r[j] = f32_add(r[j], a[i + j]);
  float s = f32_add(f32_add(f32_add(r[0], r[1]), f32_add(r[2], r[3])), f32_add(f32_add(r[4], r[5]), f32_add(r[6], r[7])));
  for (; i < n; ++i) s = f32_add(s, a[i]);
  return s;
}
CORAL_HD float np_pairwise_sum(const float* a, int n) {
  // iterative form of the recursion (explicit stack of pending right halves)
  // post-order evaluation: leaves return values that are added in the recursion's order
  struct Frame { const float* a; int n; int state; float left; };
  Frame fr[32];
  int sp = 0;
  fr[0] = Frame{a, n, 0, 0.0f};
  float ret = 0.0f;
  while (sp >= 0) {
    Frame& f = fr[sp];
    if (f.n <= 128) {
      ret = np_pairwise_leaf(f.a, f.n);
      --sp;
      continue;
    }
    int n2 = f.n / 2;
    n2 -= n2 % 8;
    if (f.state == 0) {
      f.state = 1;
      fr[sp + 1] = Frame{f.a, n2, 0, 0.0f};
      ++sp;
    } else if (f.state == 1) {
      f.left = ret;
      f.state = 2;
      fr[sp + 1] = Frame{f.a + n2, f.n - n2, 0, 0.0f};
      ++sp;
    } else {
      ret = f32_add(f.left, ret);
      --sp;
    }
  }
  return ret;
}

// Op timing for tuning: average latency of selected operations (trie find / insert, LM word
// scoring, lexicon probe, merge) as seen by the calling thread, under the real load.
#if defined(__CUDA_ARCH__)
#define CORAL_OP_T0(on) const long long _t0 = (on) ? clock64() : 0
#define CORAL_OP_T1(on, sm, slot)                                                      \
  do {                                                                                 \
    if constexpr (STATS) {                                                             \
      if (on) {                                                                        \
        atomicAdd(&(sm).tm.opc[slot], (unsigned long long)(clock64() - _t0));          \
        atomicAdd(&(sm).tm.opn[slot], 1u);                                             \
      }                                                                                \
    }                                                                                  \
  } while (0)
#else
#define CORAL_OP_T0(on) (void)0
#define CORAL_OP_T1(on, sm, slot) (void)0
#endif

// Phase timing for tuning: when a stats buffer is passed, thread 0 of the group adds the
// cycles between two marks to stats[slot] (slots 8..15). Costs nothing when stats == nullptr.
struct PhaseTimer {
  long long t;
  unsigned long long* stats;
  CORAL_DEV void start(unsigned long long* st) {
    stats = st;
#if defined(__CUDA_ARCH__)
    t = (st && threadIdx.x == 0) ? clock64() : 0;
#else
    t = 0;
#endif
  }
  CORAL_DEV void mark(int slot) {
#if defined(__CUDA_ARCH__)
    if (stats && threadIdx.x == 0) {
      const long long n = clock64();
      atomicAdd(&stats[slot], (unsigned long long)(n - t));
      t = n;
    }
#else
    (void)slot;
#endif
  }
};

// HEAVY: the instantiation contains the heavy-frame path (expand_heavy). The production launch
// runs a lean kernel without it (HEAVY = false: the hot per-frame loop keeps its register and
// instruction-cache budget) that hands utterances with many kept tokens per frame to a second,
// HEAVY kernel; the instrumented kernels and the host simulation are HEAVY only.
template <int NT, int BW, int OUTC, bool FRAMES = false, bool STATS = true, bool HEAVY = true>
struct BeamDecoder {
  using Sm = GroupShared<BW, OUTC, FRAMES, STATS, HEAVY>;
  // work counters and cycle timers exist only in the STATS instantiation: in the default one
  // this is a compile-time null and everything that hangs off it folds away
  static CORAL_DEV unsigned long long* stats_of(const UttIO& io) { return STATS ? io.stats : nullptr; }

  // candidate descriptor: representative beam, the LAST member beam (pyctcdecode keeps the
  // later candidate's tuple on a merge -- its frames), token, kind/flags
  static CORAL_DEV uint32_t info_pack(uint32_t rb, uint32_t last, uint32_t c, uint32_t kf) {
    if (FRAMES) return rb | (last << 9) | (c << 18) | (kf << 25);
    return rb | (c << 16) | (kf << 24);
  }
  static CORAL_DEV uint32_t info_rb(uint32_t i) { return FRAMES ? (i & 0x1FFu) : (i & 0xFFFFu); }
  static CORAL_DEV uint32_t info_last(uint32_t i) { return (i >> 9) & 0x1FFu; }
  static CORAL_DEV uint32_t info_c(uint32_t i) { return FRAMES ? ((i >> 18) & 0x7Fu) : ((i >> 16) & 0xFFu); }
  static CORAL_DEV uint32_t info_kf(uint32_t i) { return FRAMES ? (i >> 25) : (i >> 24); }
  static CORAL_DEV uint32_t frame_append(Sm& sm, const SlotScratch& sc, uint32_t parent, int32_t a, int32_t b) {
    if constexpr (FRAMES) {
      const uint32_t id = atom_add(&sm.wf.count, 1u);
      if (id >= sc.wf_cap) { sm.status = -4; return 0; }
      FrameRec r;
      r.parent = parent; r.start = a; r.end = b; r.pad = 0;
      sc.wf[id] = r;
      return id;
    }
    return 0;
  }
  static constexpr int HS = Sm::HS;
  // items (live prefixes x (kept tokens + 1)) above which a frame takes the heavy path
#ifndef CORAL_HEAVY_FACTOR
#define CORAL_HEAVY_FACTOR 3
#endif
  static constexpr uint32_t kHeavyItems = (uint32_t)CORAL_HEAVY_FACTOR * (uint32_t)OUTC;

  // ---- live-node hash (shared memory), keyed by the prefix hash ------------------------
  static CORAL_DEV int h_find(Sm& sm, unsigned long long key) {
    uint32_t i = (uint32_t)(key >> 20) & (HS - 1);
    for (;;) {
      const unsigned long long k = sm.hkey[i];
      if (k == key) return (int)i;
      if (k == 0) return -1;
      i = (i + 1) & (HS - 1);
    }
  }
  // returns the slot; `created` tells the caller it is the one that inserted the key
  static CORAL_DEV int h_insert(Sm& sm, unsigned long long key, bool& created) {
    uint32_t i = (uint32_t)(key >> 20) & (HS - 1);
    for (;;) {
      const unsigned long long k = atom_cas_u64(&sm.hkey[i], 0ULL, key);
      if (k == 0) { created = true; return (int)i; }
      if (k == key) { created = false; return (int)i; }
      i = (i + 1) & (HS - 1);
    }
  }

  // ---- back-pointer arena (HBM): append-only (parent index, token) records used only to
  // write out the winning transcripts; duplicates of a prefix are harmless here.
  static CORAL_DEV uint32_t arena_append(Sm& sm, const SlotScratch& sc, uint32_t parent, uint32_t tok) {
    const uint32_t id = atom_add(&sm.node_count, 1u);
    if (id >= sc.node_cap) { sm.status = -4; return 0; }
    sc.node_parent[id] = parent;
    sc.node_info[id] = tok;
    return id;
  }

  // ---- frame staging: log-softmax in float32 the way numpy evaluates it ---------------
  // x_max, tmp = x - x_max, exp, sum (numpy pairwise order for a
  // contiguous row of n <= 128: eight strided accumulators, combined as a balanced tree,
  // remainder added in order), log, tmp - log, clip to [log(1e-15), 0].
  // `rowsum` != nullptr: also record each raw row's float32 sum in numpy's order (the input
  // test of classify_rowsums), from the values already staged in shared memory.
  static CORAL_DEV_OUTLINE void stage_frames(Sm& sm, const DecodeParams& P, const UttIO& io, int t0, int nf,
                                             float* rowsum) {
    const int V = P.V;
    CORAL_LANES(NT) {
#pragma unroll 2
      // row = i / V without a runtime division: exact multiply-high for i < 2^16, V <= 64
      const uint32_t inv_v = V > 1 ? 0xFFFFFFFFu / (uint32_t)V + 1u : 0u;
      for (int i = lane; i < nf * V; i += NT) {
        const uint32_t f = V > 1 ? mulhi_u32((uint32_t)i, inv_v) : (uint32_t)i;
#if defined(__CUDA_ARCH__)
        // host-resident logits (zero-copy over PCIe): never trust a cached system-memory line
        sm.lp[f][(uint32_t)i - f * (uint32_t)V] =
            P.host_input == 1 ? __ldcv(io.logits + (size_t)t0 * V + i) : io.logits[(size_t)t0 * V + i];
#else
        sm.lp[f][(uint32_t)i - f * (uint32_t)V] = io.logits[(size_t)t0 * V + i];
#endif
      }
#if defined(__CUDA_ARCH__)
      // pull the next chunk of this utterance towards L2 while this one is decoded
      const int nxt0 = t0 + kChunk;
      if (nxt0 < io.T && P.host_input != 1) {
        const int nn = (io.T - nxt0 < kChunk ? io.T - nxt0 : kChunk) * V;
        const char* base = reinterpret_cast<const char*>(io.logits + (size_t)nxt0 * V);
        for (int off = lane * 128; off < nn * 4; off += NT * 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
      }
#endif
    }
    CORAL_GSYNC(NT);
    const float lo = -34.538776f;  // float32(log(1e-15))
    const bool as_prob = P.input_mode == 2 || (P.input_mode == 0 && sm.is_prob);
    // Staging scratch lives in the candidate arrays, which are idle between frames: one explicit
    // layout over the contiguous block o_key .. o_info (32 bytes per candidate slot).
    constexpr int kLanes8 = kChunk * 8;
    static_assert(offsetof(Sm, o_info) + sizeof(uint32_t) * OUTC - offsetof(Sm, o_key) == 32u * OUTC,
                  "candidate arrays must be contiguous");
    static_assert(32 * OUTC >= kLanes8 * (8 + 5 * 4), "staging scratch does not fit in the candidate arrays");
    unsigned char* scr = reinterpret_cast<unsigned char*>(sm.o_key);
    unsigned long long* bmask = reinterpret_cast<unsigned long long*>(scr);   // [kChunk][8] keep masks
    float* pmax = reinterpret_cast<float*>(scr + kLanes8 * 8);                // [kChunk][8] partial maxima
    float* racc = pmax + kLanes8;                                             // [kChunk][8] exp accumulators
    float* erem = racc + kLanes8;                                             // [kChunk][8] exp of the V % 8 tail
    float* bval = erem + kLanes8;                                             // [kChunk][8] best value per lane
    uint32_t* bidx = reinterpret_cast<uint32_t*>(bval + kLanes8);             // [kChunk][8] its index
    // the raw row sums are consumed (pass 2) before bval / bidx are written (argmax pass)
    float* rs8 = bval;                                                        // strided partial row sums
    float* rem8 = reinterpret_cast<float*>(bidx);                             // raw tail elements (V % 8)
    if (as_prob) {
      CORAL_LANES(NT) {
        for (int i = lane; i < nf * V; i += NT) {
          float& cell = (&sm.lp[0][0])[(i / V) * kVMax + i % V];
          float x = cell;
          x = x < 1e-15f ? 1e-15f : (x > 1.0f ? 1.0f : x);
          cell = logf(x);
        }
      }
      CORAL_GSYNC(NT);
    } else if (V < 8) {
      CORAL_LANES(NT) {
        for (int f = lane; f < nf; f += NT) {
          float* row = sm.lp[f];
          if (rowsum) {
            float r = 0.0f;
            for (int v = 0; v < V; ++v) r = f32_add(r, row[v]);
            rowsum[t0 + f] = r;
          }
          float mx = row[0];
          for (int v = 1; v < V; ++v) mx = row[v] > mx ? row[v] : mx;
          if (!isfinite(mx)) mx = 0.0f;
          float s = 0.0f;
          for (int v = 0; v < V; ++v) { row[v] = f32_add(row[v], -mx); s = f32_add(s, expf(row[v])); }
          const float ls = logf(s);
          for (int v = 0; v < V; ++v) {
            const float y = f32_add(row[v], -ls);
            row[v] = y < lo ? lo : (y > 0.0f ? 0.0f : y);
          }
        }
      }
      CORAL_GSYNC(NT);
    } else {
      // eight lanes per frame: lane j owns numpy's accumulator r[j] (elements j, j+8, ...)
      const int main_n = V - (V % 8);
      CORAL_LANES(NT) {
        for (int p = lane; p < nf * 8; p += NT) {
          const int f = p >> 3, j = p & 7;
          const float* row = sm.lp[f];
          float m = row[j];
#pragma unroll 1
          for (int i = j + 8; i < V; i += 8) m = row[i] > m ? row[i] : m;
          pmax[p] = m;
          if (rowsum) {
            float r = row[j];
#pragma unroll 1
            for (int i = j + 8; i < main_n; i += 8) r = f32_add(r, row[i]);
            rs8[p] = r;
            rem8[p] = main_n + j < V ? row[main_n + j] : 0.0f;
          }
        }
      }
      CORAL_GSYNC(NT);
      CORAL_LANES(NT) {
        for (int p = lane; p < nf * 8; p += NT) {
          const int f = p >> 3, j = p & 7;
          float* row = sm.lp[f];
          if (rowsum && j == 0) {
            const float* r = rs8 + f * 8;
            float rs = f32_add(f32_add(f32_add(r[0], r[1]), f32_add(r[2], r[3])),
                               f32_add(f32_add(r[4], r[5]), f32_add(r[6], r[7])));
#pragma unroll 1
            for (int i = main_n; i < V; ++i) rs = f32_add(rs, rem8[f * 8 + (i - main_n)]);
            rowsum[t0 + f] = rs;
          }
          float mx = pmax[f * 8];
#pragma unroll 1
          for (int k = 1; k < 8; ++k) mx = pmax[f * 8 + k] > mx ? pmax[f * 8 + k] : mx;
          if (!isfinite(mx)) mx = 0.0f;
          row[j] = f32_add(row[j], -mx);
          float r = expf(row[j]);
#pragma unroll 1
          for (int i = j + 8; i < main_n; i += 8) { row[i] = f32_add(row[i], -mx); r = f32_add(r, expf(row[i])); }
          racc[p] = r;
          if (main_n + j < V) {  // remainder: summed in order below
            row[main_n + j] = f32_add(row[main_n + j], -mx);
            erem[p] = expf(row[main_n + j]);
          }
        }
      }
      CORAL_GSYNC(NT);
      CORAL_LANES(NT) {
        for (int p = lane; p < nf * 8; p += NT) {
          const int f = p >> 3, j = p & 7;
          float* row = sm.lp[f];
          const float* r = racc + f * 8;
          float s = f32_add(f32_add(f32_add(r[0], r[1]), f32_add(r[2], r[3])),
                            f32_add(f32_add(r[4], r[5]), f32_add(r[6], r[7])));
#pragma unroll 1
          for (int i = main_n; i < V; ++i) s = f32_add(s, erem[f * 8 + (i - main_n)]);
          const float ls = logf(s);  // the same value in all eight lanes of the frame
#pragma unroll 1
          for (int i = j; i < V; i += 8) {
            const float y = f32_add(row[i], -ls);
            row[i] = y < lo ? lo : (y > 0.0f ? 0.0f : y);
          }
        }
      }
      CORAL_GSYNC(NT);
    }
    // argmax (first maximum) and the kept-token list in ascending id: eight lanes per frame
    // scan strided elements into (best value, best index, 64-bit keep mask), then one lane per
    // frame combines them and walks the mask's set bits.
    {
      CORAL_LANES(NT) {
        for (int p = lane; p < nf * 8; p += NT) {
          const int f = p >> 3, j = p & 7;
          const float* row = sm.lp[f];
          float best = -INFINITY;
          uint32_t am = 0xFFFFFFFFu;
          unsigned long long mask = 0;
#pragma unroll 1
          for (int v = j; v < V; v += 8) {
            const float x = row[v];
            if (am == 0xFFFFFFFFu || x > best) { best = x; am = (uint32_t)v; }
            if (x >= P.token_min_logp) mask |= 1ULL << v;
          }
          bval[p] = best;
          bidx[p] = am;
          bmask[p] = mask;
        }
      }
      CORAL_GSYNC(NT);
      CORAL_LANES(NT) {
        for (int f = lane; f < nf; f += NT) {
          float best = 0.0f;
          uint32_t am = 0xFFFFFFFFu;
          unsigned long long mask = 0;
#pragma unroll 1
          for (int j = 0; j < 8; ++j) {
            const uint32_t aj = bidx[f * 8 + j];
            if (aj == 0xFFFFFFFFu) continue;
            const float x = bval[f * 8 + j];
            if (am == 0xFFFFFFFFu || x > best || (x == best && aj < am)) { best = x; am = aj; }
            mask |= bmask[f * 8 + j];
          }
          mask |= 1ULL << am;
          int nk = 0;
          while (mask) {
#if defined(__CUDA_ARCH__)
            const int v = __ffsll((long long)mask) - 1;
#else
            const int v = __builtin_ctzll(mask);
#endif
            sm.kept[f][nk++] = (uint8_t)v;
            mask &= mask - 1;
          }
          sm.nkept[f] = (uint8_t)nk;
          sm.amax[f] = (uint8_t)am;
        }
      }
      CORAL_GSYNC(NT);
    }
  }

  static CORAL_DEV uint32_t rep_beam(Sm& sm, int s) { return sm.sb0[s] != kNone16 ? sm.sb0[s] : sm.sb1[s]; }

  // Sequential log-sum-exp of (logit[b] + p) over up to four member beams in ascending
  // beam index (= the reference's candidate order within one token). Returns min index.
  static CORAL_DEV_OUTLINE uint32_t merge_members(Sm& sm, int cur, uint32_t m[4], int n, double p, double& score) {
    for (int a = 1; a < n; ++a) {  // insertion sort of <= 4 indices
      uint32_t x = m[a];
      int b = a - 1;
      while (b >= 0 && m[b] > x) { m[b + 1] = m[b]; --b; }
      m[b + 1] = x;
    }
    score = d_add(sm.logit[cur][m[0]], p);
    for (int a = 1; a < n; ++a) score = sum_log_scores(score, d_add(sm.logit[cur][m[a]], p));
    return m[0];
  }

  static CORAL_DEV OutView smem_outs(Sm& sm) {
    OutView o;
    o.key = sm.o_key; o.logit = sm.o_logit; o.order = sm.o_order; o.aux = sm.o_aux; o.child = sm.o_child;
    o.info = sm.o_info;
    return o;
  }
  static CORAL_DEV double key_to_double(unsigned long long u) {
    u = (u >> 63) ? (u & 0x7FFFFFFFFFFFFFFFULL) : ~u;
    union { double d; unsigned long long u; } cv;
    cv.u = u;
    return cv.d;
  }
  // Score bucket, monotone in the score: floor((mhat - score) * scale) clamped to [0, kNB).
  // Any monotone map keeps the ranking exact; mhat only has to spread the survivors.
  static CORAL_DEV uint32_t bucket_of(double mhat, double scale, double comb) {
    const double d = d_mul(d_add(mhat, -comb), scale);
    return d >= (double)(kNB - 1) ? (uint32_t)(kNB - 1) : (d > 0.0 ? (uint32_t)d : 0u);
  }
  static CORAL_DEV double bucket_scale(const DecodeParams& P) { return P.bucket_scale; }
  // Candidates live in the shared-memory arrays while they fit (index < OUTC) and spill to the
  // slot's HBM buffer past that; the slow overflow path runs only if the frame really produced
  // more than OUTC candidates (flat logits / very wide beams).
  static CORAL_DEV_OUTLINE void emit(Sm& sm, const OutView& g, uint32_t g_cap, int q, double scale, double comb,
                                     double logit, uint32_t order, uint32_t aux, uint32_t child, uint32_t rb,
                                     uint32_t c, uint32_t kf, unsigned long long& lmax, uint32_t last) {
    const uint32_t at = atom_add(&sm.n_out[q], 1u);
    const unsigned long long k = ordered_u64(comb);
    const uint32_t info = info_pack(rb, last, c, kf);
    // counting-sort histogram for phase 3 (over every candidate: the overflow path finds its
    // cut bucket from the same counts)
    const uint32_t b = bucket_of(sm.mhat, scale, comb);
    atom_add(&sm.bcnt[b], 1u);
    if (at < (uint32_t)OUTC) {
      sm.o_key[at] = k;
      sm.o_logit[at] = logit;
      sm.o_order[at] = order;
      sm.o_aux[at] = aux;
      sm.o_child[at] = child;
      sm.o_info[at] = info;
      sm.o_bkt[at] = (uint8_t)b;
      sm.o_next[at] = (uint16_t)atom_exch_u32(&sm.bhead[b], at);
    } else if (at < g_cap) {
      g.key[at] = k;
      g.logit[at] = logit;
      g.order[at] = order;
      g.aux[at] = aux;
      g.child[at] = child;
      g.info[at] = info;
    } else {
      sm.status = -4;
    }
    lmax = k > lmax ? k : lmax;
  }

  // ---- phase 1: hash the live nodes of the current beam list -----------------------------
  // The hash was cleared during the previous frame's phase 3. The lane whose CAS inserts a
  // node registers it in the node list; every beam records itself in its node's slot.
  static CORAL_DEV void hash_beams(Sm& sm, int cur, int q, uint32_t nb, double best_lp) {
    CORAL_LANES(NT) {
      if (lane == 0) {
        // bucket reference for this frame: last frame's best score + this frame's best
        // log-prob (+2: merges and word completions can raise a score a little)
        sm.mhat = d_add(d_add(key_to_double(sm.gmax[q ^ 1]), best_lp), 2.0);
      }
      for (int i = lane; i < kNB; i += NT) { sm.bcnt[i] = 0; sm.bhead[i] = kListEnd; }
      for (uint32_t b = lane; b < nb; b += NT) {
        const uint32_t mt = sm.meta[cur][b];
        bool created;
        const int s = h_insert(sm, sm.nh[cur][b], created);
        if (created) sm.ne_slot[atom_add(&sm.nN[q], 1u)] = (uint16_t)s;
        if (meta_lc(mt) == kLcBlank) sm.sb0[s] = (uint16_t)b; else sm.sb1[s] = (uint16_t)b;
      }
    }
    CORAL_GSYNC(NT);
  }

  static CORAL_DEV void clear_hash(Sm& sm, int lane) {
    for (int i = lane; i < HS; i += NT) { sm.hkey[i] = 0; sm.sb0[i] = kNone16; sm.sb1[i] = kNone16; }
  }

  static CORAL_DEV void expand_item(Sm& sm, const LmView& lm, const DecodeParams& P, const SlotScratch& sc,
                                    const UttIO& io, int f, int cur, int q, uint32_t nb, const OutView& outs,
                                    double bscale, int K, bool fam2, uint32_t j, uint32_t k_in,
                                    unsigned long long& lmax) {
    {
      {
#if defined(__CUDA_ARCH__)
        const bool tim = stats_of(io) != nullptr;  // tuning: where a letter item spends its cycles
        const long long tA = tim ? clock64() : 0;
        long long tB = 0, tC = 0, tD = 0;
#endif
        const int s = sm.ne_slot[j];
        const uint32_t rb = rep_beam(sm, s);
        const uint32_t mt = sm.meta[cur][rb];
        const uint32_t tok_m = meta_tok(mt), fl_m = meta_flags(mt), wlen_m = sm.wlen[cur][rb];
        const uint32_t b0 = sm.sb0[s], b1 = sm.sb1[s];
        uint32_t mem[4];
        int nm = 0;
        double logit;
        if (fam2) {
          // repeat of the node's own last token (or a space on a closed word): the text does
          // not change. If the parent node is live, its (parent, c) item gathers these beams.
          const uint32_t c = tok_m == kNoTok ? (uint32_t)P.space_id : tok_m;
          if (tok_m != kNoTok && h_find(sm, sm.ph[cur][rb]) >= 0) return;
          int k = -1;
          for (int kk = 0; kk < K; ++kk) if (sm.kept[f][kk] == c) k = kk;
          if (k < 0) return;
          if (b1 != kNone16) mem[nm++] = b1;  // last_char == c (None or space at the root)
          if ((int)c == P.space_id && b0 != kNone16) mem[nm++] = b0;
          if (nm == 0) return;
          const uint32_t first = merge_members(sm, cur, mem, nm, (double)sm.lp[f][c], logit);
          emit(sm, outs, sc.outs_cap, q, bscale, d_add(logit, d_add(sm.lm_raw[cur][rb], partial_score(P, wlen_m, fl_m))), logit,
               (uint32_t)k * nb + first, 0u, 0u, rb, c, 0u, lmax, mem[nm - 1]);
          return;
        }
        const uint32_t k = k_in;
        const uint32_t c = sm.kept[f][k];
        const double p = (double)sm.lp[f][c];
        if ((int)c == P.blank_id) {
          if (b0 != kNone16) mem[nm++] = b0;
          if (b1 != kNone16) mem[nm++] = b1;
          const uint32_t first = merge_members(sm, cur, mem, nm, p, logit);
          emit(sm, outs, sc.outs_cap, q, bscale, d_add(logit, d_add(sm.lm_raw[cur][rb], partial_score(P, wlen_m, fl_m))), logit,
               k * nb + first, 0u, 0u, rb, c, 0u, lmax, mem[nm - 1]);
          return;
        }
        if ((int)c == P.space_id && wlen_m == 0) return;  // a space after a closed word never extends
        if (b0 != kNone16) mem[nm++] = b0;
        if (b1 != kNone16 && tok_m != c) mem[nm++] = b1;
        const int cs = h_find(sm, child_hash(sm.nh[cur][rb], c));
        if (cs >= 0) {
          if (sm.sb1[cs] != kNone16) mem[nm++] = sm.sb1[cs];
          if ((int)c == P.space_id && sm.sb0[cs] != kNone16) mem[nm++] = sm.sb0[cs];
        }
        if (nm == 0) return;
#if defined(__CUDA_ARCH__)
        if (tim) tB = clock64();
#endif
        const uint32_t first = merge_members(sm, cur, mem, nm, p, logit);
        const uint32_t order = k * nb + first;
#if defined(__CUDA_ARCH__)
        if (tim) tC = clock64();
#endif
        if (cs >= 0) {
          const uint32_t crb = rep_beam(sm, cs);
          emit(sm, outs, sc.outs_cap, q, bscale,
               d_add(logit, d_add(sm.lm_raw[cur][crb],
                                  partial_score(P, sm.wlen[cur][crb], meta_flags(sm.meta[cur][crb])))),
               logit, order, 0u, crb, rb, c, 1u, lmax, mem[nm - 1]);
        } else if ((int)c == P.space_id) {
          // a word closes: score it with the LM and keep the result in a boundary record
          // (what pyctcdecode caches under the new text in cached_lm_scores)
          uint32_t bnd_new = 0;
          double raw_new = sm.lm_raw[cur][rb];
          if (lm.present || P.prune_history) {
            bnd_new = atom_add(&sm.bnd_count, 1u);
            if (bnd_new >= sc.bnd_cap) { sm.status = -4; bnd_new = 0; }
          }
          if (lm.present) {
            const bool in_lm = (fl_m & kInLm) != 0;
            const bool oov = (lm.has_unigrams && !(fl_m & kInUni)) || !in_lm;
            BndRec nr;
            CORAL_OP_T0(stats_of(io) != nullptr);
            const double sw = lm_word_score(lm, P, sc.bnd[sm.bnd[cur][rb]].st, in_lm ? sm.wid[cur][rb] : 0u, oov,
                                            false, nr.st, stats_of(io) ? sm.cnt : nullptr);
            CORAL_OP_T1(stats_of(io) != nullptr, sm, 1);
            nr.lm_raw = d_add(sm.lm_raw[cur][rb], sw);
            raw_new = nr.lm_raw;
            sc.bnd[bnd_new] = nr;
          }
          emit(sm, outs, sc.outs_cap, q, bscale, d_add(logit, d_add(raw_new, 0.0)), logit, order, bnd_new, 0u, rb, c, 3u, lmax, mem[nm - 1]);
        } else {
          // a letter extends the partial word: roll the word hash, probe the lexicon
          uint32_t nfl = 0, nwid = 0;
          double ps = 0.0;
          if (lm.present) {
            if (fl_m & kDead) {
              nfl = kDead | kOovPartial;
            } else {
              unsigned long long h = sm.whash[cur][rb];
              for (int qq = 0; qq < P.label_ncp[c]; ++qq) h = word_hash_push(h, P.label_cps[c][qq]);
              uint32_t lw, lf;
              if (stats_of(io)) atom_add(&sm.cnt[4], 1u);
              bool lfound;
              { CORAL_OP_T0(stats_of(io) != nullptr); lfound = lex_find(lm, h, lw, lf); CORAL_OP_T1(stats_of(io) != nullptr, sm, 3); }
              if (lfound) {
                nfl = ((lf & kLexPrefixOfUnigram) ? 0u : kOovPartial) | ((lf & kLexInUnigrams) ? kInUni : 0u) |
                      ((lf & kLexInLm) ? kInLm : 0u);
                nwid = lw;
              } else {
                nfl = kDead | kOovPartial;
              }
            }
            ps = partial_score(P, wlen_m + P.label_ncp[c], nfl);
          }
#if defined(__CUDA_ARCH__)
          if (tim) tD = clock64();
#endif
          emit(sm, outs, sc.outs_cap, q, bscale, d_add(logit, d_add(sm.lm_raw[cur][rb], ps)), logit, order, nwid, 0u, rb, c,
               2u | (nfl << 2), lmax, mem[nm - 1]);
#if defined(__CUDA_ARCH__)
          if constexpr (STATS) {
            if (tim) {
              const long long tE = clock64();
              atomicAdd(&sm.tm.opc[0], (unsigned long long)(tB - tA)); atomicAdd(&sm.tm.opn[0], 1u);
              atomicAdd(&sm.tm.opc[2], (unsigned long long)(tC - tB));
              atomicAdd(&sm.tm.opc[4], (unsigned long long)(tD - tC));
              atomicAdd(&sm.tm.opc[7], (unsigned long long)(tE - tD));
            }
          }
#endif
        }
      }
    }
  }

  // ---- phase 2: every (live node, kept token), plus repeats of nodes whose parent is dead ---
  static CORAL_DEV void expand(Sm& sm, const LmView& lm, const DecodeParams& P, const SlotScratch& sc,
                               const UttIO& io, int f, int cur, int q, uint32_t nb, const OutView& outs) {
    const int K = sm.nkept[f];
    const double bscale = bucket_scale(P);
    const uint32_t nN = sm.nN[q];
    CORAL_LANES(NT) {
      if (lane == 0) {
        // clear the other parity's per-frame counters for the next frame. This must sit behind
        // this frame's first barrier: every thread read S[q ^ 1] (the beam count) on its way in.
        sm.nN[q ^ 1] = 0; sm.n_out[q ^ 1] = 0; sm.S[q ^ 1] = 0; sm.gmax[q ^ 1] = 0;
        if (stats_of(io)) { sm.cnt[0] += (uint32_t)K * nb; sm.cnt[3] += 1u; }
      }
      // Work units = (token group g, chunk of 32 nodes); group K is the "repeat" family. A warp
      // takes whole units, so its lanes follow the same code path (same token kind) instead of
      // serialising the blank / space / letter / repeat paths inside every warp, and the groups
      // of a typical frame (K + 1 <= 4) run side by side on the CTA's warps.
      unsigned long long lmax = 0;
      const uint32_t nchunks = (nN + 31u) >> 5;
      const uint32_t units = ((uint32_t)K + 1u) * nchunks;
      constexpr uint32_t kWarps = NT / 32;
#pragma unroll 1
      uint32_t g = 0, ch = (uint32_t)lane >> 5;  // unit u = g * nchunks + ch, advanced without dividing
      while (nchunks && ch >= nchunks) { ch -= nchunks; ++g; }
      for (uint32_t u = (uint32_t)lane >> 5; u < units; u += kWarps) {
        const uint32_t j = ch * 32u + ((uint32_t)lane & 31u);
        if (j < nN) expand_item(sm, lm, P, sc, io, f, cur, q, nb, outs, bscale, K, g == (uint32_t)K, j, g, lmax);
        ch += kWarps;
        while (ch >= nchunks) { ch -= nchunks; ++g; }
      }
      if (lmax) atom_max_u64(&sm.gmax[q], lmax);
    }
    CORAL_GSYNC(NT);
  }

  // ---- heavy frames: many kept tokens (flat posteriors, loose token_min_logp) ----------------
  // nN x (K + 1) items, of which only beam_width can survive. Instead of expanding all of them,
  // the items are visited in the order of a cheap UPPER BOUND of their score (the threshold
  // algorithm of top-k query processing):
  //   A. every item: members through the live-prefix hash (as in expand_item), upper bound
  //      ub = max member logit + ln(#members) + p + an upper bound of the LM part (exact for
  //      blank / repeat / merge items; "no partial-word penalty" for a letter; alpha * (unk) *
  //      ln10 + beta for a word that closes) -> histogram of upper-bound bins;
  //   B. counting sort of the items by bin (best first) into the slot's HBM scratch;
  //   C. chunks of NT items in that order go through expand_item (the exact, expensive work:
  //      lexicon probe, LM scoring, log-sum-exp). After each chunk: as soon as beam_width exact
  //      candidates sit in buckets strictly better than the best remaining upper bound -- or that
  //      bound is below the prune threshold -- no remaining item can survive, and the frame is done.
  // Exact: ub >= true score item by item, the bucket map is monotone, and ranking / pruning only
  // ever look at candidates that could make the cut.
  static constexpr uint32_t kNoBin = 0xFFu;
  static_assert(HS <= 1024, "SlotScratch::hv_masks holds 1024 slots");
  // What the bound of an item needs to know about its live prefix, fetched once per prefix.
  struct HeavyNode {
    double l0, l1;        // logit of the member ending in a blank / in the prefix's last token (-inf: none)
    double u_ab, u_a;     // max + ln(n) + 1e-6 over {b0, b1} and over {b0} alone (-inf: no member)
    double own;           // LM part of the unchanged text: lm_raw + partial-word score
    double lm_raw, pen1;  // raw LM score; partial-word penalty of one more (one-code-point) letter
    unsigned long long ok, child_live;
    uint32_t rb, tok_m, fl_m, wlen_m, b0, b1;
    int s;
    bool ok_known, parent_live;
  };
  static CORAL_DEV void heavy_node_setup(Sm& sm, const LmView& lm, const DecodeParams& P, const SlotScratch& sc,
                                         int cur, uint32_t j, HeavyNode& n) {
    n.s = sm.ne_slot[j];
    n.rb = rep_beam(sm, n.s);
    const uint32_t mt = sm.meta[cur][n.rb];
    n.tok_m = meta_tok(mt); n.fl_m = meta_flags(mt); n.wlen_m = sm.wlen[cur][n.rb];
    n.b0 = sm.sb0[n.s]; n.b1 = sm.sb1[n.s];
    n.l0 = n.b0 != kNone16 ? sm.logit[cur][n.b0] : -INFINITY;
    n.l1 = n.b1 != kNone16 ? sm.logit[cur][n.b1] : -INFINITY;
    // + 1e-6: the sequential fp64 log-sum-exp may round a hair above max + ln(n)
    n.u_a = n.l0 + 1e-6;
    n.u_ab = (n.l0 > n.l1 ? n.l0 : n.l1) + ((n.b0 != kNone16 && n.b1 != kNone16) ? 0.6931471805599454 : 0.0) + 1e-6;
    n.lm_raw = sm.lm_raw[cur][n.rb];
    n.own = d_add(n.lm_raw, partial_score(P, n.wlen_m, n.fl_m));
    n.pen1 = lm.present ? partial_score(P, n.wlen_m + 1u, kOovPartial) : 0.0;
    // tokens whose label keeps the partial word penalty-free (exact, from the lexicon's child mask)
    n.ok = ~0ULL;
    n.ok_known = false;
    if (lm.present && lm.lex_ok != nullptr && !(n.fl_m & kDead)) {
      n.ok_known = true;
      if (n.wlen_m == 0) n.ok = lm.root_ok;
      else {
        const long long sl = lex_slot(lm, sm.whash[cur][n.rb]);
        n.ok = sl >= 0 ? lm.lex_ok[sl] : 0ULL;
      }
    }
    n.parent_live = n.tok_m != kNoTok && h_find(sm, sm.ph[cur][n.rb]) >= 0;
    n.child_live = sc.hv_masks()[n.s];
  }
  // the rare item kinds: repeat family, space, a letter whose extended prefix is live already
  static CORAL_DEV_OUTLINE double heavy_ub_general(Sm& sm, const LmView& lm, const DecodeParams& P, const HeavyNode& n,
                                                   int f, int cur, int K, bool fam2, uint32_t c) {
    static const double kLn[5] = {0.0, 0.0, 0.6931471805599454, 1.0986122886681098, 1.3862943611198906};
    int nm = 0;
    double mx = -INFINITY, lm_ub = n.own;
    if (fam2) {
      if (n.parent_live) return -INFINITY;
      bool kept = false;
      for (int kk = 0; kk < K; ++kk) kept |= sm.kept[f][kk] == c;
      if (!kept) return -INFINITY;
      if (n.b1 != kNone16) { ++nm; mx = n.l1; }
      if ((int)c == P.space_id && n.b0 != kNone16) { ++nm; mx = n.l0 > mx ? n.l0 : mx; }
    } else {
      if ((int)c == P.space_id && n.wlen_m == 0) return -INFINITY;
      if (n.b0 != kNone16) { ++nm; mx = n.l0; }
      if (n.b1 != kNone16 && n.tok_m != c) { ++nm; mx = n.l1 > mx ? n.l1 : mx; }
      const int cs = ((n.child_live >> c) & 1ULL) ? h_find(sm, child_hash(sm.nh[cur][n.rb], c)) : -1;
      if (cs >= 0) {
        const uint32_t c1 = sm.sb1[cs], c0 = sm.sb0[cs];
        if (c1 != kNone16) { ++nm; const double v = sm.logit[cur][c1]; mx = v > mx ? v : mx; }
        if ((int)c == P.space_id && c0 != kNone16) { ++nm; const double v = sm.logit[cur][c0]; mx = v > mx ? v : mx; }
        const uint32_t crb = rep_beam(sm, cs);
        lm_ub = d_add(sm.lm_raw[cur][crb], partial_score(P, sm.wlen[cur][crb], meta_flags(sm.meta[cur][crb])));
      } else if ((int)c == P.space_id) {
        lm_ub = n.lm_raw;
        if (lm.present) {
          if (P.alpha < 0.0) return INFINITY;  // no upper bound on alpha * log p: visit it first
          const bool oov = (lm.has_unigrams && !(n.fl_m & kInUni)) || !(n.fl_m & kInLm);
          // the word scores at most alpha * (log10 p upper bound + unk offset if OOV) * ln10 + beta
          const double x = d_add((double)lm.score_ub, oov ? P.unk_score_offset : 0.0);
          lm_ub = d_add(lm_ub, d_add(d_mul(d_mul(P.alpha, x), P.log_base_change), P.beta));
        }
      } else {
        lm_ub = n.lm_raw;
        if (lm.present) {
          const bool penal = (n.fl_m & kDead) || (n.ok_known && !((n.ok >> c) & 1ULL));
          if (penal || (!n.ok_known && P.unk_score_offset > 0.0))
            lm_ub = d_add(lm_ub, partial_score(P, n.wlen_m + P.label_ncp[c], kOovPartial));
        }
      }
    }
    if (nm == 0) return -INFINITY;
    return mx + kLn[nm] + (double)sm.lp[f][c] + lm_ub + 1e-6;
  }
  // upper-bound bin of item (prefix n, token group g); kNoBin = the item produces no candidate
  static CORAL_DEV uint32_t heavy_bin(Sm& sm, const LmView& lm, const DecodeParams& P, const HeavyNode& n, int f,
                                      int cur, int K, uint32_t g, double bscale) {
    double ub;
    if (g == (uint32_t)K) {
      ub = heavy_ub_general(sm, lm, P, n, f, cur, K, true, n.tok_m == kNoTok ? (uint32_t)P.space_id : n.tok_m);
    } else {
      const uint32_t c = sm.kept[f][g];
      if ((int)c == P.blank_id) {
        ub = n.u_ab + (double)sm.lp[f][c] + n.own;                     // -inf when the prefix has no member at all
      } else if ((int)c == P.space_id || ((n.child_live >> c) & 1ULL) || P.label_ncp[c] != 1) {
        ub = heavy_ub_general(sm, lm, P, n, f, cur, K, false, c);
      } else {
        // a plain letter onto a prefix that is not live yet: members b0 (+ b1 unless it repeats c)
        const double u = n.tok_m == c ? n.u_a : n.u_ab;
        const bool penal = lm.present && ((n.fl_m & kDead) || (n.ok_known && !((n.ok >> c) & 1ULL)) ||
                                          (!n.ok_known && P.unk_score_offset > 0.0));
        ub = u + (double)sm.lp[f][c] + n.lm_raw + (penal ? n.pen1 : 0.0);
      }
    }
    if (!(ub > -INFINITY)) return kNoBin;
    return bucket_of(sm.mhat, bscale, ub) >> 1;
  }

  static CORAL_DEV_OUTLINE void expand_heavy(Sm& sm, const LmView& lm, const DecodeParams& P, const SlotScratch& sc,
                                             const UttIO& io, int f, int cur, int q, uint32_t nb, const OutView& outs) {
    const int K = sm.nkept[f];
    const double bscale = bucket_scale(P);
    const uint32_t nN = sm.nN[q];
    constexpr int kBins = kNB / 2;
    uint32_t* sorted = sc.hv_sorted();
    CORAL_LANES(NT) {
      if (lane == 0) {
        sm.nN[q ^ 1] = 0; sm.n_out[q ^ 1] = 0; sm.S[q ^ 1] = 0; sm.gmax[q ^ 1] = 0;
        if (stats_of(io)) { sm.cnt[0] += (uint32_t)K * nb; sm.cnt[3] += 1u; }
      }
      for (int i = lane; i < kBins; i += NT) sm.hv.ubcnt[i] = 0;
      for (int i = lane; i < HS; i += NT) sc.hv_masks()[i] = 0ULL;
    }
    CORAL_GSYNC(NT);
    // A0. every live prefix tells its (live) parent which token leads to it: nN hash look-ups
    // instead of one per (prefix, kept token)
    CORAL_LANES(NT) {
      for (uint32_t j = (uint32_t)lane; j < nN; j += NT) {
        const uint32_t rb = rep_beam(sm, sm.ne_slot[j]);
        const uint32_t tok = meta_tok(sm.meta[cur][rb]);
        if (tok == kNoTok) continue;
        const int sp = h_find(sm, sm.ph[cur][rb]);
        if (sp >= 0) atom_or_u64(&sc.hv_masks()[sp], 1ULL << tok);
      }
    }
    CORAL_GSYNC(NT);
    // A. upper-bound bin of every item (one live prefix per thread), kept for the scatter below
    uint8_t* bins = sc.hv_bin();
    const uint32_t M = ((uint32_t)K + 1u) * nN;
    CORAL_LANES(NT) {
      for (uint32_t j = (uint32_t)lane; j < nN; j += NT) {
        HeavyNode n;
        heavy_node_setup(sm, lm, P, sc, cur, j, n);
        for (uint32_t g = 0; g <= (uint32_t)K; ++g) {
          const uint32_t bin = heavy_bin(sm, lm, P, n, f, cur, K, g, bscale);
          bins[g * nN + j] = (uint8_t)bin;
          if (bin != kNoBin) atom_add_u16(&sm.hv.ubcnt[bin], 1u);
        }
      }
    }
    CORAL_GSYNC(NT);
    // B. bin counts -> start offsets (warp 0: two bins per lane + a shuffle scan), then the item ids
    // in the order of their bins
    CORAL_LANES(NT) {
#if defined(__CUDA_ARCH__)
      static_assert(kBins == 64, "two upper-bound bins per lane");
      if (lane < 32) {
        const uint32_t c0 = sm.hv.ubcnt[2 * lane], c1 = sm.hv.ubcnt[2 * lane + 1];
        uint32_t incl = c0 + c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        const uint32_t excl = incl - (c0 + c1);
        sm.hv.ubcnt[2 * lane] = (uint16_t)excl;
        sm.hv.ubcnt[2 * lane + 1] = (uint16_t)(excl + c0);
        if (lane == 31) sm.sel_n = incl;  // items that produce a candidate at all
      }
#else
      if (lane == 0) {
        uint32_t run = 0;
        for (int b = 0; b < kBins; ++b) { const uint32_t c = sm.hv.ubcnt[b]; sm.hv.ubcnt[b] = (uint16_t)run; run += c; }
        sm.sel_n = run;
      }
#endif
    }
    CORAL_GSYNC(NT);
    const uint32_t Mv = sm.sel_n;
    CORAL_LANES(NT) {
      for (uint32_t it = (uint32_t)lane; it < M; it += NT) {
        const uint32_t bin = bins[it];
        if (bin != kNoBin) sorted[atom_add_u16(&sm.hv.ubcnt[bin], 1u)] = (bin << 24) | it;
      }
    }
    CORAL_GSYNC(NT);
    // C. exact expansion in upper-bound order, one chunk of NT items at a time, until nothing that
    // is left can survive
    for (uint32_t base = 0; base < Mv; base += NT) {
      CORAL_LANES(NT) {
        unsigned long long lmax = 0;
        const uint32_t at = base + (uint32_t)lane;
        if (at < Mv) {
          const uint32_t it = sorted[at] & 0xFFFFFFu;
          const uint32_t g = it / nN, j = it - g * nN;
          expand_item(sm, lm, P, sc, io, f, cur, q, nb, outs, bscale, K, g == (uint32_t)K, j, g, lmax);
        }
        if (lmax) atom_max_u64(&sm.gmax[q], lmax);
      }
      CORAL_GSYNC(NT);
      if (sm.status != 0) return;
      const uint32_t next = base + NT;
      if (next >= Mv) break;
      // the best upper bound still unvisited, as an exact-bucket index and as a score
      const uint32_t nbkt = (sorted[next] >> 24) * 2u;
      const double ub_score = d_add(sm.mhat, -d_div((double)nbkt, bscale));  // every score in bucket >= nbkt is <= this
      bool done = ordered_u64(ub_score) < prune_key(sm, P, q) && nbkt > 0;
      if (!done) {
        uint32_t cum = 0;
        for (uint32_t b = 0; b < nbkt && b < (uint32_t)kNB; ++b) cum += sm.bcnt[b];
        done = cum >= (uint32_t)P.beam_width;
      }
      if (done) break;
    }
    CORAL_GSYNC(NT);
  }

  // ---- overflow path: more candidates than the shared-memory arrays hold ------------------
  // Every candidate was counted into its score bucket while it was produced, so the bucket that
  // holds the beam_width-th best score is known: the candidates of the buckets up to that one
  // are pulled from HBM into shared memory (one pass) and phase 3 ranks them exactly as in the
  // common case. Only if those buckets hold more than the shared arrays (scores bunched in one
  // bucket) a radix select over the 64-bit keys picks the beam_width best. `thr` = prune key.
  static CORAL_DEV_OUTLINE void select_overflow(Sm& sm, const DecodeParams& P, const OutView& g, int q,
                                        unsigned long long thr) {
    const uint32_t n = sm.n_out[q];
    CORAL_LANES(NT) {
      for (uint32_t i = lane; i < (uint32_t)OUTC; i += NT) {
        g.key[i] = sm.o_key[i]; g.logit[i] = sm.o_logit[i]; g.order[i] = sm.o_order[i];
        g.aux[i] = sm.o_aux[i]; g.child[i] = sm.o_child[i]; g.info[i] = sm.o_info[i];
      }
      if (lane == 0) {
        uint32_t cum = 0;
        int b = 0;
        for (; b < kNB - 1; ++b) { cum += sm.bcnt[b]; if (cum >= (uint32_t)P.beam_width) break; }
        if (b == kNB - 1) cum += sm.bcnt[b];
        sm.sel_cut = (uint32_t)b;
        sm.sel_eq = cum;  // candidates in the buckets up to the cut
        sm.sel_n = 0;
      }
    }
    CORAL_GSYNC(NT);
    if (sm.sel_eq > (uint32_t)OUTC) { select_overflow_radix(sm, P, g, q, thr); return; }
    const uint32_t cutb = sm.sel_cut;
    const double scale = bucket_scale(P);
    CORAL_LANES(NT) {
      for (uint32_t base = lane; base < n; base += 4 * NT) {
        unsigned long long k[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const uint32_t i = base + u * NT; k[u] = i < n ? g.key[i] : 0ULL; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t i = base + u * NT;
          if (i >= n || k[u] < thr) continue;
          if (bucket_of(sm.mhat, scale, key_to_double(k[u])) > cutb) continue;
          const uint32_t at = atom_add(&sm.sel_n, 1u);  // < OUTC: at most sel_eq candidates qualify
          sm.o_key[at] = k[u]; sm.o_logit[at] = g.logit[i]; sm.o_order[at] = g.order[i]; sm.o_aux[at] = g.aux[i];
          sm.o_child[at] = g.child[i]; sm.o_info[at] = g.info[i];
        }
      }
    }
    CORAL_GSYNC(NT);
    CORAL_LANES(NT) { if (lane == 0) sm.n_out[q] = sm.sel_n; }
    CORAL_GSYNC(NT);
  }
  static CORAL_DEV_OUTLINE void select_overflow_radix(Sm& sm, const DecodeParams& P, const OutView& g, int q,
                                              unsigned long long thr) {
    const uint32_t n = sm.n_out[q];
    CORAL_LANES(NT) { if (lane == 0) sm.cnt[7] += 1u; }  // counter: frames that needed the radix select
    uint32_t* hist = reinterpret_cast<uint32_t*>(sm.o_key);  // 256 bins; o_key is free until the pull
    CORAL_LANES(NT) { if (lane == 0) { sm.sel_prefix = 0; sm.sel_mask = 0; sm.sel_need = (uint32_t)P.beam_width; sm.sel_n = 0; } }
    CORAL_GSYNC(NT);
    for (int pass = 0; pass < 8; ++pass) {
      const int shift = 56 - 8 * pass;
      CORAL_LANES(NT) { for (int i = lane; i < 256; i += NT) hist[i] = 0; }
      CORAL_GSYNC(NT);
      const unsigned long long pre = sm.sel_prefix, msk = sm.sel_mask;
      CORAL_LANES(NT) {
        for (uint32_t i = lane; i < n; i += NT) {
          const unsigned long long k = g.key[i];
          if (k >= thr && (k & msk) == pre) atom_add(&hist[(k >> shift) & 255], 1u);
        }
      }
      CORAL_GSYNC(NT);
      CORAL_LANES(NT) {
        for (int gq = lane; gq < 16; gq += NT) {
          uint32_t t = 0;
          for (int d = 0; d < 16; ++d) t += hist[gq * 16 + d];
          sm.gsum[gq] = t;
        }
      }
      CORAL_GSYNC(NT);
      CORAL_LANES(NT) {
        if (lane == 0) {
          uint32_t cum = 0, need = sm.sel_need;
          int gq = 15;
          for (; gq > 0; --gq) { if (cum + sm.gsum[gq] >= need) break; cum += sm.gsum[gq]; }
          int d = gq * 16 + 15;
          for (; d > gq * 16; --d) { if (cum + hist[d] >= need) break; cum += hist[d]; }
          sm.sel_need = need - cum;
          sm.sel_eq = hist[d];
          sm.sel_prefix = pre | ((unsigned long long)d << shift);
          sm.sel_mask = msk | (255ULL << shift);
        }
      }
      CORAL_GSYNC(NT);
    }
    const unsigned long long kth = sm.sel_prefix;
    // fewer than beam_width survivors in total: everything >= thr is taken (kth <= thr then)
    uint32_t ord_cut = 0xFFFFFFFFu;
    if (sm.sel_eq > sm.sel_need) {  // ties at the cut: keep the sel_need smallest candidate indices
      CORAL_LANES(NT) {
        if (lane == 0) {
          uint32_t last = 0, cut = 0;
          bool first = true;
          for (uint32_t r = 0; r < sm.sel_need; ++r) {
            uint32_t best = 0xFFFFFFFFu;
            for (uint32_t i = 0; i < n; ++i)
              if (g.key[i] == kth && (first || g.order[i] > last) && g.order[i] < best) best = g.order[i];
            last = best;
            first = false;
            cut = best;
          }
          sm.sel_cut = cut;
        }
      }
      CORAL_GSYNC(NT);
      ord_cut = sm.sel_cut;
    }
    CORAL_LANES(NT) {
      for (uint32_t i = lane; i < n; i += NT) {
        const unsigned long long k = g.key[i];
        if (k < thr) continue;
        if (k > kth || (k == kth && g.order[i] <= ord_cut)) {
          const uint32_t at = atom_add(&sm.sel_n, 1u);
          if (at < (uint32_t)OUTC) {
            sm.o_logit[at] = g.logit[i]; sm.o_order[at] = g.order[i]; sm.o_aux[at] = g.aux[i];
            sm.o_child[at] = g.child[i]; sm.o_info[at] = g.info[i];
            g.order[i] |= 0x80000000u;  // mark: its key is copied after the histogram area is free
            g.aux[i] = at;
          }
        }
      }
    }
    CORAL_GSYNC(NT);
    CORAL_LANES(NT) {
      for (uint32_t i = lane; i < n; i += NT)
        if (g.order[i] & 0x80000000u) sm.o_key[g.aux[i]] = g.key[i];
      if (lane == 0) sm.n_out[q] = sm.sel_n < (uint32_t)OUTC ? sm.sel_n : (uint32_t)OUTC;
    }
    CORAL_GSYNC(NT);
  }

  // ---- phase 3: prune, trim to beam_width, rank, write the next beam list -------------------
  // rank = number of candidates that beat this one under (score desc, first-candidate index
  // asc) -- pyctcdecode's stable heapq.nlargest. Candidates were dropped into kNB score
  // buckets (monotone in the score) while they were produced, so
  //   rank = (candidates in better buckets) + (bucket-mates that beat it):
  // a counting sort by bucket, then comparisons inside the bucket only -- ~100 instructions
  // per survivor instead of a pass over every candidate, exact for any score distribution.
  // Candidates below the prune threshold never beat a survivor, so they need no filter.
  static CORAL_DEV void rebucket(Sm& sm, const DecodeParams& P, int q) {  // overflow path only
    const uint32_t n = sm.n_out[q];
    const double scale = bucket_scale(P);
    CORAL_LANES(NT) {
      for (int i = lane; i < kNB; i += NT) { sm.bcnt[i] = 0; sm.bhead[i] = kListEnd; }
    }
    CORAL_GSYNC(NT);
    CORAL_LANES(NT) {
      for (uint32_t i = lane; i < n; i += NT) {
        const uint32_t b = bucket_of(sm.mhat, scale, key_to_double(sm.o_key[i]));
        atom_add(&sm.bcnt[b], 1u);
        sm.o_bkt[i] = (uint8_t)b;
        sm.o_next[i] = (uint16_t)atom_exch_u32(&sm.bhead[b], i);
      }
    }
    CORAL_GSYNC(NT);
  }
  // Exclusive prefix sums of the bucket counts -> bstart. Every warp computes the same 128 values
  // and stores them (identical stores, so no block barrier is needed: a warp reads only after its
  // own stores). Host simulation: lane 0 does it sequentially.
  static CORAL_DEV void scan_buckets(Sm& sm, int lane) {
#if defined(__CUDA_ARCH__)
    static_assert(kNB % 32 == 0, "bucket count must be a multiple of the warp size");
    constexpr int PER = kNB / 32;
    const int l = lane & 31;
    uint32_t c[PER];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) { c[k] = sm.bcnt[l * PER + k]; s += c[k]; }
    uint32_t incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (l >= o) incl += v;
    }
    uint32_t run = incl - s;
#pragma unroll
    for (int k = 0; k < PER; ++k) { sm.bstart[l * PER + k] = (uint16_t)(run > 0xFFFFu ? 0xFFFFu : run); run += c[k]; }
#else
    if (lane == 0) {
      uint32_t run = 0;
      for (int b = 0; b < kNB; ++b) { sm.bstart[b] = (uint16_t)(run > 0xFFFFu ? 0xFFFFu : run); run += sm.bcnt[b]; }
    }
#endif
  }
  static CORAL_DEV void rank_and_commit(Sm& sm, const LmView& lm, const DecodeParams& P, const SlotScratch& sc,
                                        int cur, int q, unsigned long long thr, bool final_pass,
                                        PhaseTimer* pt = nullptr, int t = 0) {
    const int nxt = cur ^ 1;
    const uint32_t n = sm.n_out[q];
    CORAL_LANES(NT) { scan_buckets(sm, lane); }
    CORAL_WSYNC();
    if (pt) pt->mark(12);
    CORAL_LANES(NT) {
      if (!final_pass) clear_hash(sm, lane);
      for (uint32_t i = lane; i < n; i += NT) {
        const unsigned long long ki = sm.o_key[i];
        if (ki < thr) continue;
        const uint32_t b = sm.o_bkt[i];
        uint32_t r = sm.bstart[b];
        // beam_width or more candidates sit in strictly better buckets: it cannot survive the trim
        if (r >= (uint32_t)P.beam_width) continue;
        const uint32_t oi = sm.o_order[i];
#pragma unroll 2
        for (uint32_t j = sm.bhead[b]; j != kListEnd; j = sm.o_next[j]) {
          const unsigned long long kj = sm.o_key[j];
          r += kj > ki;
          if (kj == ki) r += sm.o_order[j] < oi;
        }
        if (r >= (uint32_t)P.beam_width) continue;
        atom_add(&sm.S[q], 1u);
        const double logit = sm.o_logit[i];
        if constexpr (FRAMES) {
          // word timing of the member pyctcdecode keeps (the later candidate), SURVEY A5 step 4
          const uint32_t inf = sm.o_info[i];
          const uint32_t L = info_last(inf), c = info_c(inf);
          int32_t p0 = sm.wf.pf0[cur][L], p1 = sm.wf.pf1[cur][L];
          uint32_t hd = sm.wf.head[cur][L];
          if (final_pass) {
            if (sm.wlen[cur][L] > 0) hd = frame_append(sm, sc, hd, p0, p1);
          } else if ((int)c == P.blank_id) {
          } else if (c == meta_lc(sm.meta[cur][L])) {
            p1 = t + 1;
          } else if ((int)c == P.space_id) {
            if (sm.wlen[cur][L] > 0) hd = frame_append(sm, sc, hd, p0, p1);
            p0 = p1 = -1;
          } else {
            if (p0 < 0) p0 = t;
            p1 = t + 1;
          }
          sm.wf.pf0[nxt][r] = p0; sm.wf.pf1[nxt][r] = p1; sm.wf.head[nxt][r] = hd;
        }
        if (final_pass) {
          const double comb = key_to_double(ki);
          sm.logit[nxt][r] = logit;
          sm.lm_raw[nxt][r] = comb;        // final: combined score travels in lm_raw
          sm.node[nxt][r] = sm.o_child[i];  // final: text node
          continue;
        }
        const uint32_t info = sm.o_info[i];
        const uint32_t rb = info_rb(info), c = info_c(info), kf = info_kf(info);
        const uint32_t kind = kf & 3u;
        const uint32_t lc = (int)c == P.blank_id ? kLcBlank : c;
        if (kind == 0 || kind == 1) {
          const uint32_t src = kind == 0 ? rb : sm.o_child[i];
          const uint32_t mt = sm.meta[cur][src];
          sm.node[nxt][r] = sm.node[cur][src];
          sm.nh[nxt][r] = sm.nh[cur][src];
          sm.ph[nxt][r] = sm.ph[cur][src];
          sm.bnd[nxt][r] = sm.bnd[cur][src];
          sm.wid[nxt][r] = sm.wid[cur][src];
          sm.whash[nxt][r] = sm.whash[cur][src];
          sm.lm_raw[nxt][r] = sm.lm_raw[cur][src];
          sm.wlen[nxt][r] = sm.wlen[cur][src];
          sm.meta[nxt][r] = meta_pack(meta_tok(mt), lc, meta_flags(mt));
        } else {
          // a new prefix: append its back-pointer (stores only, nothing waits on HBM here)
          sm.node[nxt][r] = arena_append(sm, sc, sm.node[cur][rb], c);
          sm.nh[nxt][r] = child_hash(sm.nh[cur][rb], c);
          sm.ph[nxt][r] = sm.nh[cur][rb];
          if (kind == 2) {
            unsigned long long h = sm.whash[cur][rb];
            for (int qq = 0; qq < P.label_ncp[c]; ++qq) h = word_hash_push(h, P.label_cps[c][qq]);
            const uint32_t wl = (uint32_t)sm.wlen[cur][rb] + P.label_ncp[c];
            sm.bnd[nxt][r] = sm.bnd[cur][rb];
            sm.wid[nxt][r] = sm.o_aux[i];
            sm.whash[nxt][r] = h;
            sm.lm_raw[nxt][r] = sm.lm_raw[cur][rb];
            sm.wlen[nxt][r] = (uint16_t)(wl > 65535u ? 65535u : wl);
            sm.meta[nxt][r] = meta_pack(c, lc, kf >> 2);
          } else {
            const uint32_t bn = sm.o_aux[i];
            if (P.prune_history) hist_push(sc, sm.bnd[cur][rb], bn, sm.whash[cur][rb], P.prune_history);
            sm.bnd[nxt][r] = bn;
            sm.wid[nxt][r] = 0;
            sm.whash[nxt][r] = kWordHashSeed;
            sm.lm_raw[nxt][r] = lm.present ? sc.bnd[bn].lm_raw : 0.0;
            sm.wlen[nxt][r] = 0;
            sm.meta[nxt][r] = meta_pack(c, lc, 0u);
          }
        }
        sm.logit[nxt][r] = logit;
      }
    }
    CORAL_GSYNC(NT);
    if (pt) pt->mark(13);
  }

  // ---- pyctcdecode prune_history=True (UP:pyctcdecode decoder.py _prune_history) ----
  static CORAL_DEV_OUTLINE void hist_push(const SlotScratch& sc, uint32_t parent, uint32_t rec, unsigned long long word,
                                          int n) {
    HistRec h;
    const HistRec& p = sc.hist()[parent];
    h.w[0] = word ? word : 1ULL;
    for (int k = 1; k < kMaxCtx; ++k) h.w[k] = p.w[k - 1];
    unsigned long long x = 0x6A09E667F3BCC909ULL;
    for (int k = 0; k < n && k < kMaxCtx; ++k) x = mix64(x ^ h.w[k]) + 0x9E3779B97F4A7C15ULL;
    h.H = x;
    sc.hist()[rec] = h;
  }
  // After the trim, in rank order: keep the first beam of every (last n words of the text,
  // word_part, last_char). The beam list is compacted in place (read everything, barrier, write).
  static CORAL_DEV_OUTLINE void prune_history_pass(Sm& sm, const DecodeParams& P, const SlotScratch& sc, int nxt, int q) {
    const uint32_t S = sm.S[q] < (uint32_t)P.beam_width ? sm.S[q] : (uint32_t)P.beam_width;
    CORAL_LANES(NT) {
      for (uint32_t r = lane; r < S; r += NT) {
        const uint32_t mt = sm.meta[nxt][r];
        unsigned long long k = sc.hist()[sm.bnd[nxt][r]].H;
        k = mix64(k ^ sm.whash[nxt][r]) + (unsigned long long)sm.wlen[nxt][r] * 0x9E3779B97F4A7C15ULL;
        k = mix64(k ^ (unsigned long long)meta_lc(mt));
        sm.o_key[r] = k;
      }
    }
    CORAL_GSYNC(NT);
    CORAL_LANES(NT) {
      for (uint32_t r = lane; r < S; r += NT) {
        const unsigned long long k = sm.o_key[r];
        uint32_t dup = 0;
        for (uint32_t j = 0; j < r; ++j) dup |= sm.o_key[j] == k ? 1u : 0u;
        sm.o_order[r] = dup ? 0u : 1u;
      }
    }
    CORAL_GSYNC(NT);
    // every lane handles the beams r = lane, lane + NT, ...: at most (BW + NT - 1) / NT of them
    constexpr int PER = (BW + NT - 1) / NT;
    double v_logit[PER], v_raw[PER];
    unsigned long long v_wh[PER], v_nh[PER], v_ph[PER];
    uint32_t v_node[PER], v_bnd[PER], v_wid[PER], v_meta[PER], v_dst[PER];
    uint16_t v_wlen[PER];
    int32_t v_p0[PER], v_p1[PER];
    uint32_t v_hd[PER];
#if defined(CORAL_HOSTSIM)
    // lanes run one after the other on the host: stage through per-beam copies instead of registers
    static thread_local double h_logit[BW], h_raw[BW];
    static thread_local unsigned long long h_wh[BW], h_nh[BW], h_ph[BW];
    static thread_local uint32_t h_node[BW], h_bnd[BW], h_wid[BW], h_meta[BW], h_dst[BW], h_hd[BW];
    static thread_local uint16_t h_wlen[BW];
    static thread_local int32_t h_p0[BW], h_p1[BW];
    uint32_t kept = 0;
    for (uint32_t r = 0; r < S; ++r) {
      h_dst[r] = sm.o_order[r] ? kept++ : kNoNode;
      h_logit[r] = sm.logit[nxt][r]; h_raw[r] = sm.lm_raw[nxt][r]; h_wh[r] = sm.whash[nxt][r];
      h_nh[r] = sm.nh[nxt][r]; h_ph[r] = sm.ph[nxt][r]; h_node[r] = sm.node[nxt][r]; h_bnd[r] = sm.bnd[nxt][r];
      h_wid[r] = sm.wid[nxt][r]; h_meta[r] = sm.meta[nxt][r]; h_wlen[r] = sm.wlen[nxt][r];
      if constexpr (FRAMES) { h_p0[r] = sm.wf.pf0[nxt][r]; h_p1[r] = sm.wf.pf1[nxt][r]; h_hd[r] = sm.wf.head[nxt][r]; }
    }
    for (uint32_t r = 0; r < S; ++r) {
      const uint32_t d = h_dst[r];
      if (d == kNoNode) continue;
      sm.logit[nxt][d] = h_logit[r]; sm.lm_raw[nxt][d] = h_raw[r]; sm.whash[nxt][d] = h_wh[r];
      sm.nh[nxt][d] = h_nh[r]; sm.ph[nxt][d] = h_ph[r]; sm.node[nxt][d] = h_node[r]; sm.bnd[nxt][d] = h_bnd[r];
      sm.wid[nxt][d] = h_wid[r]; sm.meta[nxt][d] = h_meta[r]; sm.wlen[nxt][d] = h_wlen[r];
      if constexpr (FRAMES) { sm.wf.pf0[nxt][d] = h_p0[r]; sm.wf.pf1[nxt][d] = h_p1[r]; sm.wf.head[nxt][d] = h_hd[r]; }
    }
    sm.S[q] = kept;
    (void)v_logit; (void)v_raw; (void)v_wh; (void)v_nh; (void)v_ph; (void)v_node; (void)v_bnd; (void)v_wid;
    (void)v_meta; (void)v_dst; (void)v_wlen; (void)v_p0; (void)v_p1; (void)v_hd;
#else
    {
      const int lane = (int)(threadIdx.x % NT);
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        const uint32_t r = (uint32_t)lane + (uint32_t)u * NT;
        v_dst[u] = kNoNode;
        if (r < S && sm.o_order[r]) {
          uint32_t d = 0;
          for (uint32_t j = 0; j < r; ++j) d += sm.o_order[j];
          v_dst[u] = d;
          v_logit[u] = sm.logit[nxt][r]; v_raw[u] = sm.lm_raw[nxt][r]; v_wh[u] = sm.whash[nxt][r];
          v_nh[u] = sm.nh[nxt][r]; v_ph[u] = sm.ph[nxt][r]; v_node[u] = sm.node[nxt][r]; v_bnd[u] = sm.bnd[nxt][r];
          v_wid[u] = sm.wid[nxt][r]; v_meta[u] = sm.meta[nxt][r]; v_wlen[u] = sm.wlen[nxt][r];
          if constexpr (FRAMES) { v_p0[u] = sm.wf.pf0[nxt][r]; v_p1[u] = sm.wf.pf1[nxt][r]; v_hd[u] = sm.wf.head[nxt][r]; }
        }
      }
      CORAL_GSYNC(NT);
      uint32_t mine = 0;
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        const uint32_t d = v_dst[u];
        if (d == kNoNode) continue;
        ++mine;
        sm.logit[nxt][d] = v_logit[u]; sm.lm_raw[nxt][d] = v_raw[u]; sm.whash[nxt][d] = v_wh[u];
        sm.nh[nxt][d] = v_nh[u]; sm.ph[nxt][d] = v_ph[u]; sm.node[nxt][d] = v_node[u]; sm.bnd[nxt][d] = v_bnd[u];
        sm.wid[nxt][d] = v_wid[u]; sm.meta[nxt][d] = v_meta[u]; sm.wlen[nxt][d] = v_wlen[u];
        if constexpr (FRAMES) { sm.wf.pf0[nxt][d] = v_p0[u]; sm.wf.pf1[nxt][d] = v_p1[u]; sm.wf.head[nxt][d] = v_hd[u]; }
      }
      if (lane == 0) sm.S[q] = 0;
      CORAL_GSYNC(NT);
      if (mine) atom_add(&sm.S[q], mine);
    }
#endif
    CORAL_GSYNC(NT);
  }

  static CORAL_DEV unsigned long long prune_key(Sm& sm, const DecodeParams& P, int q) {
    // max_score + beam_prune_logp in float64, back in key space
    return ordered_u64(d_add(key_to_double(sm.gmax[q]), P.beam_prune_logp));
  }

  // ---- one frame: 3 barriers on the common path ------------------------------------------------
  static CORAL_DEV void frame_step(Sm& sm, const LmView& lm, const DecodeParams& P, SlotScratch& sc,
                                   const UttIO& io, int f, int cur, int q, uint32_t nb, int t) {
    PhaseTimer pt;
    pt.start(stats_of(io));
    hash_beams(sm, cur, q, nb, (double)sm.lp[f][sm.amax[f]]);
    pt.mark(8);
#if defined(__CUDA_ARCH__)
    const uint32_t bnd_before = sm.bnd_count;
    const long long t_exp = stats_of(io) ? clock64() : 0;
#endif
    // frames with many kept tokens go through the threshold algorithm (flat posteriors, loose
    // token_min_logp); the common frames (K + 1 <= 4 on trained-model-like posteriors) never do
    if constexpr (HEAVY) {
      if (((uint32_t)sm.nkept[f] + 1u) * sm.nN[q] > kHeavyItems)
        expand_heavy(sm, lm, P, sc, io, f, cur, q, nb, sc.outs_g);
      else
        expand(sm, lm, P, sc, io, f, cur, q, nb, sc.outs_g);
    } else {
      expand(sm, lm, P, sc, io, f, cur, q, nb, sc.outs_g);
    }
    pt.mark(9);
#if defined(__CUDA_ARCH__)
    if (stats_of(io) && threadIdx.x == 0) {  // tuning: expand time split by "frame scored a word with the LM"
      const int slot = sm.bnd_count != bnd_before ? 5 : 6;
      if constexpr (STATS) {
        sm.tm.opc[slot] += (unsigned long long)(clock64() - t_exp);
        sm.tm.opn[slot] += 1u;
      }
    }
#endif
    if (sm.status != 0) return;  // uniform: written before the barrier that ends expand
    const unsigned long long thr = prune_key(sm, P, q);
    if (sm.n_out[q] > (uint32_t)OUTC) {
      select_overflow(sm, P, sc.outs_g, q, thr);
      rebucket(sm, P, q);
      pt.mark(10);
    }
    rank_and_commit(sm, lm, P, sc, cur, q, thr, false, &pt, t);
    if (P.prune_history && sm.status == 0) prune_history_pass(sm, P, sc, cur ^ 1, q);
  }

  // ---- end of utterance ------------------------------------------------------
  static CORAL_DEV_OUTLINE uint32_t finalize(Sm& sm, const LmView& lm, const DecodeParams& P, const SlotScratch& sc,
                                     const UttIO& io, int cur, int q, uint32_t nb) {
    hash_beams(sm, cur, q, nb, 0.0);
    const uint32_t nN = sm.nN[q];
    const OutView outs = sc.outs_g;  // never reached: at most beam_width <= OUTC candidates here
    const double bscale = bucket_scale(P);
    CORAL_LANES(NT) {
      unsigned long long lmax = 0;
      for (uint32_t j = lane; j < nN; j += NT) {
        const int s = sm.ne_slot[j];
        const uint32_t rb = rep_beam(sm, s);
        const uint32_t mt = sm.meta[cur][rb];
        const bool open_or_root = sm.wlen[cur][rb] > 0 || meta_tok(mt) == kNoTok;
        // a closed-word node whose (open-word) parent is live is absorbed by the parent's group
        if (!open_or_root && h_find(sm, sm.ph[cur][rb]) >= 0) continue;
        uint32_t mem[4];
        int nm = 0;
        if (sm.sb0[s] != kNone16) mem[nm++] = sm.sb0[s];
        if (sm.sb1[s] != kNone16) mem[nm++] = sm.sb1[s];
        if (sm.wlen[cur][rb] > 0) {
          const int cs = h_find(sm, child_hash(sm.nh[cur][rb], (uint32_t)P.space_id));
          if (cs >= 0) {
            if (sm.sb0[cs] != kNone16) mem[nm++] = sm.sb0[cs];
            if (sm.sb1[cs] != kNone16) mem[nm++] = sm.sb1[cs];
          }
        }
        double logit;
        // logit + 0.0 is exact, so merge_members' "+ p" leaves the scores untouched
        const uint32_t first = merge_members(sm, cur, mem, nm, 0.0, logit);
        const uint32_t last = mem[nm - 1];  // the later candidate's tuple is the one pyctcdecode keeps
        double comb = logit;
        if (lm.present) {
          const bool open = sm.wlen[cur][last] > 0;
          const uint32_t lfl = meta_flags(sm.meta[cur][last]);
          const bool in_lm = open && (lfl & kInLm);
          const bool oov = !open || (lm.has_unigrams && !(lfl & kInUni)) || !in_lm;
          LmState out;
          const double sw = lm_word_score(lm, P, sc.bnd[sm.bnd[cur][last]].st, in_lm ? sm.wid[cur][last] : 0u, oov,
                                          true, out, stats_of(io) ? sm.cnt : nullptr);
          comb = d_add(logit, d_add(d_add(sm.lm_raw[cur][last], sw), 0.0));
        }
        // text node: the open-word node itself, or the parent of a closed-word node
        emit(sm, outs, sc.outs_cap, q, bscale, comb, logit, first, 0u,
             open_or_root ? sm.node[cur][rb] : sc.node_parent[sm.node[cur][rb]], rb, 0u, 0u, lmax, last);
      }
      if (lmax) atom_max_u64(&sm.gmax[q], lmax);
    }
    CORAL_GSYNC(NT);
    rank_and_commit(sm, lm, P, sc, cur, q, prune_key(sm, P, q), true);
    const int fin = cur ^ 1;
    const uint32_t nf = sm.S[q] < (uint32_t)P.beam_width ? sm.S[q] : (uint32_t)P.beam_width;
    CORAL_LANES(NT) {
      if (lane == 0) {
        *io.out_n = (int32_t)nf;
        *io.out_status = sm.status;
        if (stats_of(io)) {
          sm.cnt[5] = sm.node_count;
          sm.cnt[6] = sm.bnd_count;
          for (int k = 0; k < 8; ++k) atom_add(&stats_of(io)[k], (unsigned long long)sm.cnt[k]);
          if constexpr (STATS) {
            for (int k = 0; k < 8; ++k) { atom_add(&stats_of(io)[16 + k], sm.tm.opc[k]); atom_add(&stats_of(io)[24 + k], (unsigned long long)sm.tm.opn[k]); }
          }
        }
      }
      for (uint32_t r = lane; r < nf && r < (uint32_t)P.n_best; r += NT) {
        io.out_logit[r] = sm.logit[fin][r];
        io.out_comb[r] = sm.lm_raw[fin][r];
        uint8_t* dst = io.out_tokens + (size_t)r * P.T_max;
        uint32_t n = sm.node[fin][r];
        int len = 0;
        while (n != 0 && len < P.T_max) {
          dst[len++] = (uint8_t)(sc.node_info[n] & 0xFFu);
          n = sc.node_parent[n];
        }
        for (int a = 0, b = len - 1; a < b; ++a, --b) { const uint8_t t = dst[a]; dst[a] = dst[b]; dst[b] = t; }
        io.out_len[r] = len;
        if constexpr (FRAMES) {
          // text_frames of the final beam, oldest word first
          int nw = 0;
          for (uint32_t h = sm.wf.head[fin][r]; h != 0; h = sc.wf[h].parent) ++nw;
          int32_t* fr = io.out_frames + (size_t)r * io.max_words * 2;
          int w = nw - 1;
          for (uint32_t h = sm.wf.head[fin][r]; h != 0; h = sc.wf[h].parent, --w) {
            if (w < io.max_words) { fr[2 * w] = sc.wf[h].start; fr[2 * w + 1] = sc.wf[h].end; }
          }
          io.out_nwords[r] = nw < io.max_words ? nw : io.max_words;
        }
      }
    }
    CORAL_GSYNC(NT);
    return nf;
  }

  // ---- pyctcdecode's probabilities-vs-logits test for one utterance -> sm.is_prob.
  // math.isclose(logits.sum(axis=1).mean(), 1) holds only when the float32 mean is exactly
  // 1.0f. The row sums were recorded in numpy's order while the frames were staged (no second
  // pass over the logits). Their mean is first bounded with a double-precision total: unless
  // it lies within 1e-3 of 1 the float32 mean cannot be 1.0f whatever the summation order
  // (float32 pairwise error is < 1e-5 relative), and only then one lane replays numpy's
  // pairwise order over the T row sums.
  static CORAL_DEV_OUTLINE void classify_rowsums(Sm& sm, const SlotScratch& sc, int T) {
    static_assert(OUTC >= NT, "the candidate arrays double as per-lane scratch");
    double* dsum = reinterpret_cast<double*>(sm.o_logit);  // [NT], candidate arrays are idle here
    double* dabs = reinterpret_cast<double*>(sm.o_key);    // [NT]
    CORAL_LANES(NT) {
      double ds = 0.0, da = 0.0;
      for (int f = lane; f < T; f += NT) {
        const float s = sc.rowsum[f];
        ds += (double)s;
        da += (double)(s < 0.0f ? -s : s);
      }
      dsum[lane] = ds;
      dabs[lane] = da;
    }
    CORAL_GSYNC(NT);
    CORAL_LANES(NT) {
      if (lane == 0) {
        int p = 0;
        if (T > 0) {
          double ds = 0.0, da = 0.0;
          for (int k = 0; k < NT; ++k) { ds += dsum[k]; da += dabs[k]; }
          const double dev = ds / T - 1.0;
          if ((dev < 0 ? -dev : dev) <= 1e-3 * (1.0 + da / T)) {  // false for NaN: not probabilities
            const float tot = np_pairwise_sum(sc.rowsum, T);
#if defined(__CUDA_ARCH__)
            const float mean = __fdiv_rn(tot, (float)T);
#else
            const float mean = tot / (float)T;
#endif
            p = mean == 1.0f ? 1 : 0;
          }
        }
        sm.is_prob = p;
      }
    }
    CORAL_GSYNC(NT);
  }

  // ---- whole utterance ---------------------------------------------------------------------------
  // With input_mode 0 (pyctcdecode's auto-detection) the utterance is decoded as logits while
  // the row sums are recorded; in the rare case that they then say "probabilities", it is
  // decoded again as such. Real callers pass logits, so nothing is read or done twice.
  static CORAL_DEV void decode(Sm& sm, const LmView& lm, const DecodeParams& P, SlotScratch& sc, const UttIO& io) {
    int cur = 0, q = 0;
    uint32_t nb = 1;
#pragma unroll 1
    for (int attempt = 0; attempt < 2; ++attempt) {
    const bool speculate = P.input_mode == 0 && attempt == 0;
    CORAL_LANES(NT) {
      clear_hash(sm, lane);
      if (lane == 0 && attempt == 0) sm.is_prob = 0;
      if (lane == 0) {
        sm.status = 0;
        for (int k = 0; k < 8; ++k) sm.cnt[k] = 0;
        if constexpr (STATS) { for (int k = 0; k < 8; ++k) { sm.tm.opc[k] = 0; sm.tm.opn[k] = 0; } }
        sm.node_count = 1;
        sm.bnd_count = 1;
        sm.nN[0] = sm.nN[1] = 0;
        sm.n_out[0] = sm.n_out[1] = 0;
        sm.S[0] = sm.S[1] = 0;
        sm.gmax[0] = 0;
        sm.gmax[1] = ordered_u64(0.0);  // "previous frame's best score" of the empty beam
        sm.logit[0][0] = 0.0;
        sm.lm_raw[0][0] = 0.0;
        sm.whash[0][0] = kWordHashSeed;
        sm.node[0][0] = 0;
        sm.nh[0][0] = kRootHash;
        sm.ph[0][0] = 0;
        sm.bnd[0][0] = 0;
        sm.wid[0][0] = 0;
        sm.wlen[0][0] = 0;
        sm.meta[0][0] = meta_pack(kNoTok, kLcNone, 0u);
        if constexpr (FRAMES) { sm.wf.pf0[0][0] = -1; sm.wf.pf1[0][0] = -1; sm.wf.head[0][0] = 0; sm.wf.count = 1; }
        sc.node_parent[0] = kNoNode;
        sc.node_info[0] = kNoTok;
        if (P.prune_history) {
          HistRec h0;
          h0.H = 0x6A09E667F3BCC909ULL;
          for (int k = 0; k < kMaxCtx; ++k) h0.w[k] = 0;
          // the root's H must be what hist_push would produce from an empty window
          unsigned long long x = 0x6A09E667F3BCC909ULL;
          for (int k = 0; k < P.prune_history && k < kMaxCtx; ++k) x = mix64(x ^ 0ULL) + 0x9E3779B97F4A7C15ULL;
          h0.H = x;
          sc.hist()[0] = h0;
        }
        if (lm.present) {
          BndRec r0;
          r0.lm_raw = 0.0;
          if (P.score_boundary) lm_begin_sentence(lm, r0.st); else lm_null_context(r0.st);
          sc.bnd[0] = r0;
        }
      }
    }
    CORAL_GSYNC(NT);
    cur = 0; q = 0; nb = 1;
    bool failed = false;
    for (int t0 = 0; t0 < io.T && !failed; t0 += kChunk) {
      const int nf = io.T - t0 < kChunk ? io.T - t0 : kChunk;
      {
        PhaseTimer ps;
        ps.start(stats_of(io));
        stage_frames(sm, P, io, t0, nf, speculate ? sc.rowsum : nullptr);
        ps.mark(15);
      }
      for (int f = 0; f < nf; ++f) {
        frame_step(sm, lm, P, sc, io, f, cur, q, nb, t0 + f);
        if (sm.status != 0) { failed = true; break; }
        cur ^= 1;
        nb = sm.S[q] < (uint32_t)P.beam_width ? sm.S[q] : (uint32_t)P.beam_width;
        q ^= 1;
      }
    }
    if (failed) {
      CORAL_LANES(NT) { if (lane == 0) { *io.out_n = 0; *io.out_status = sm.status; } }
      CORAL_GSYNC(NT);
      return;
    }
    if (!speculate) break;
    classify_rowsums(sm, sc, io.T);
    if (!sm.is_prob) break;
    }  // attempt
    finalize(sm, lm, P, sc, io, cur, q, nb);
  }
};

}  // namespace coral
