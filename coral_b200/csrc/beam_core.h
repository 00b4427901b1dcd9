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
#include <math.h>
#include <stdint.h>

#include "lm_tables.h"

namespace coral {

#if defined(CORAL_HOSTSIM)
#define CORAL_DEV inline
template <class T>
inline T atom_add(T* p, T v) { T o = *p; *p = o + v; return o; }
inline unsigned long long atom_max_u64(unsigned long long* p, unsigned long long v) {
  unsigned long long o = *p; if (v > o) *p = v; return o;
}
inline unsigned long long atom_cas_u64(unsigned long long* p, unsigned long long c, unsigned long long v) {
  unsigned long long o = *p; if (o == c) *p = v; return o;
}
#define CORAL_LANES(NT) for (int lane = 0; lane < (NT); ++lane)
#define CORAL_GSYNC(NT) ((void)0)
#else
#define CORAL_DEV __device__ __forceinline__
template <class T>
__device__ __forceinline__ T atom_add(T* p, T v) { return atomicAdd(p, v); }
__device__ __forceinline__ unsigned long long atom_max_u64(unsigned long long* p, unsigned long long v) {
  return atomicMax(p, v);
}
__device__ __forceinline__ unsigned long long atom_cas_u64(unsigned long long* p, unsigned long long c,
                                                           unsigned long long v) {
  return atomicCAS(p, c, v);
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
constexpr int kChunk = 16;       // frames staged in shared memory per pass
constexpr int kMaxLabelCps = 8;  // code points per alphabet label
constexpr uint32_t kNone16 = 0xFFFFu;
constexpr uint32_t kNoTok = 0xFFu;
constexpr uint32_t kNoNode = 0xFFFFFFFFu;

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
  int32_t input_mode;   // 0 auto (is_prob[] decides), 1 logits, 2 probabilities
  int32_t score_boundary;
  float token_min_logp;  // compared in float32
  double beam_prune_logp;
  double alpha, beta, unk_score_offset;
  double log_base_change;  // 1 / log10(e) as the reference computes it
  uint32_t label_cps[kVMax][kMaxLabelCps];
  uint8_t label_ncp[kVMax];
};

struct BndRec {  // LM record of a word boundary (a complete-words text prefix)
  double lm_raw;
  LmState st;
};

struct alignas(8) OutRec {
  double comb;
  double logit;
  uint32_t order;  // first candidate index in (token-major, beam-minor) order
  uint32_t aux;    // kind 2: LM word id ; kind 3: boundary record
  uint32_t child;  // kind 1: representative beam of the live child ; kind 3: node id
  uint16_t slot;   // hash slot of the source node
  uint8_t c;       // token
  uint8_t kf;      // kind (low 2 bits) | lexicon flags << 2
};

struct SlotScratch {  // per thread-group arenas in HBM, reused utterance after utterance
  uint32_t* node_parent;
  uint32_t* node_info;  // tok | bnd << 8
  unsigned long long* ch_keys;  // epoch << 32 | parent << 8 | tok ; 0 = never used
  uint32_t* ch_vals;
  BndRec* bnd;
  OutRec* outs_g;       // overflow for frames with more candidates than fit in smem
  uint16_t* surv_g;
  uint32_t node_cap, bnd_cap, ch_mask, outs_cap;
  uint32_t epoch;
};

struct UttIO {
  const float* logits;  // [T, V] of this utterance
  int32_t T;
  int32_t is_prob;
  // outputs
  int32_t* out_n;       // scalar: number of final beams
  double* out_logit;    // [n_best]
  double* out_comb;     // [n_best]
  uint8_t* out_tokens;  // [n_best, T_max]
  int32_t* out_len;     // [n_best]
  int32_t* out_status;  // scalar: 0 ok, -4 capacity
  unsigned long long* stats;  // optional [8]: extensions, LM scorings, n-gram probes, frames, lexicon probes
};

template <int BW, int OUTC>
struct GroupShared {
  static constexpr int HS = 2 * BW;
  // beams, double buffered
  double logit[2][BW];
  double lm_raw[2][BW];
  unsigned long long whash[2][BW];
  uint32_t node[2][BW];
  uint32_t parent[2][BW];
  uint32_t bnd[2][BW];
  uint32_t wid[2][BW];
  uint32_t meta[2][BW];  // tok | lc << 8 | flags << 16 | wlen << 20
  // live-node hash
  unsigned long long hkey[HS];
  uint16_t sb0[HS], sb1[HS];
  uint16_t ne_slot[BW];
  uint16_t beam_slot[BW];
  uint8_t claimed[HS];
  // candidates of this frame
  OutRec outs[OUTC];
  uint16_t surv[OUTC];
  uint32_t hist[256];
  // staged frames
  float lp[kChunk][kVMax];
  uint8_t kept[kChunk][kVMax];
  uint8_t nkept[kChunk];
  uint8_t krank[kVMax];
  // scalars
  unsigned long long gmax;
  unsigned long long sel_prefix, sel_mask;
  uint32_t nb, nN, n_out, S, S2;
  uint32_t node_count, bnd_count;
  uint32_t sel_need, sel_eq;
  int32_t cur, status, utt;
};

CORAL_HD uint32_t meta_pack(uint32_t tok, uint32_t lc, uint32_t flags, uint32_t wlen) {
  return tok | (lc << 8) | (flags << 16) | ((wlen > 4095u ? 4095u : wlen) << 20);
}
CORAL_HD uint32_t meta_tok(uint32_t m) { return m & 0xFFu; }
CORAL_HD uint32_t meta_lc(uint32_t m) { return (m >> 8) & 0xFFu; }
CORAL_HD uint32_t meta_flags(uint32_t m) { return (m >> 16) & 0xFu; }
CORAL_HD uint32_t meta_wlen(uint32_t m) { return m >> 20; }

CORAL_HD unsigned long long node_key(uint32_t parent, uint32_t tok) {
  // root is (kNoNode, kNoTok); +1 keeps every key non-zero
  return (((unsigned long long)parent << 8) | tok) + 1ULL;
}

// pyctcdecode LanguageModel.score_partial_token, hotwords empty
CORAL_HD double partial_score(const DecodeParams& P, uint32_t wlen, uint32_t flags) {
  if (wlen == 0) return 0.0;
  double u = d_mul(P.unk_score_offset, (flags & kOovPartial) ? 1.0 : 0.0);
  if (wlen > 6) u = d_div(d_mul(u, (double)wlen), 6.0);
  return u;
}

// pyctcdecode LanguageModel.score: alpha * log10-score * ln10 + beta
CORAL_DEV double lm_word_score(const LmView& lm, const DecodeParams& P, const LmState& in, uint32_t wid,
                              bool oov, bool is_last, LmState& out, unsigned long long* stats) {
  int np = 0;
  double x = (double)lm_base_score(lm, in, wid, out, &np);
  if (oov) x = d_add(x, P.unk_score_offset);
  if (is_last) {
    double e = 0.0;
    if (P.score_boundary) {
      LmState tmp;
      int np2 = 0;
      e = (double)lm_base_score(lm, out, lm.eos_id, tmp, &np2);
      np += np2;
    }
    x = d_add(x, e);
  }
  if (stats) { atom_add(&stats[1], 1ULL); atom_add(&stats[2], (unsigned long long)np); }
  return d_add(d_mul(d_mul(P.alpha, x), P.log_base_change), P.beta);
}

template <int NT, int BW, int OUTC>
struct BeamDecoder {
  using Sm = GroupShared<BW, OUTC>;
  static constexpr int HS = Sm::HS;

  // ---- live-node hash (shared memory) ------------------------------------------------
  static CORAL_DEV int h_find(Sm& sm, unsigned long long key) {
    uint32_t i = (uint32_t)mix64(key) & (HS - 1);
    for (;;) {
      const unsigned long long k = sm.hkey[i];
      if (k == key) return (int)i;
      if (k == 0) return -1;
      i = (i + 1) & (HS - 1);
    }
  }
  static CORAL_DEV int h_insert(Sm& sm, unsigned long long key) {
    uint32_t i = (uint32_t)mix64(key) & (HS - 1);
    for (;;) {
      const unsigned long long k = atom_cas_u64(&sm.hkey[i], 0ULL, key);
      if (k == 0 || k == key) return (int)i;
      i = (i + 1) & (HS - 1);
    }
  }

  // ---- per-utterance trie child table (HBM) ------------------------------------------
  // Returns the node id of (parent, tok), creating it (with info word `info`) if absent.
  static CORAL_DEV uint32_t trie_get_or_add(Sm& sm, const SlotScratch& sc, uint32_t parent, uint32_t tok,
                                            uint32_t bnd, bool& created) {
    const unsigned long long tag = (unsigned long long)sc.epoch << 32;
    const unsigned long long key = tag | ((unsigned long long)(parent & 0xFFFFFFu) << 8) | tok;
    uint32_t i = (uint32_t)mix64(key) & sc.ch_mask;
    created = false;
    uint32_t id = kNoNode;
    for (;;) {
      unsigned long long k = sc.ch_keys[i];
      if (k == key) return sc.ch_vals[i];
      if ((k >> 32) != sc.epoch) {  // empty or left over from an earlier utterance
        if (id == kNoNode) id = atom_add(&sm.node_count, 1u);
        if (id >= sc.node_cap) { sm.status = -4; return 0; }
        const unsigned long long old = atom_cas_u64(&sc.ch_keys[i], k, key);
        if (old == k) {
          sc.ch_vals[i] = id;
          sc.node_parent[id] = parent;
          sc.node_info[id] = tok | (bnd << 8);
          created = true;
          return id;
        }
        // lost the slot to another lane (a different key: keys are unique per frame)
        continue;
      }
      i = (i + 1) & sc.ch_mask;
    }
  }
  static CORAL_DEV bool trie_find(const SlotScratch& sc, uint32_t parent, uint32_t tok, uint32_t& id) {
    const unsigned long long tag = (unsigned long long)sc.epoch << 32;
    const unsigned long long key = tag | ((unsigned long long)(parent & 0xFFFFFFu) << 8) | tok;
    uint32_t i = (uint32_t)mix64(key) & sc.ch_mask;
    for (;;) {
      const unsigned long long k = sc.ch_keys[i];
      if (k == key) { id = sc.ch_vals[i]; return true; }
      if ((k >> 32) != sc.epoch) return false;
      i = (i + 1) & sc.ch_mask;
    }
  }

  // ---- frame staging: log-softmax in float32 the way numpy evaluates it ---------------
  // x_max, tmp = x - x_max, exp, sum (numpy pairwise order for a
  // contiguous row of n <= 128: eight strided accumulators, combined as a balanced tree,
  // remainder added in order), log, tmp - log, clip to [log(1e-15), 0].
  static CORAL_DEV void stage_frames(Sm& sm, const DecodeParams& P, const UttIO& io, int t0, int nf) {
    const int V = P.V;
    CORAL_LANES(NT) {
      for (int i = lane; i < nf * V; i += NT) sm.lp[i / V][i % V] = io.logits[(size_t)t0 * V + i];
    }
    CORAL_GSYNC(NT);
    const float lo = -34.538776f;  // float32(log(1e-15))
    const bool as_prob = P.input_mode == 2 || (P.input_mode == 0 && io.is_prob);
    CORAL_LANES(NT) {
      for (int f = lane; f < nf; f += NT) {
        float* row = sm.lp[f];
        if (as_prob) {
          for (int v = 0; v < V; ++v) {
            float x = row[v];
            x = x < 1e-15f ? 1e-15f : (x > 1.0f ? 1.0f : x);
            row[v] = logf(x);
          }
        } else {
          float mx = row[0];
          for (int v = 1; v < V; ++v) mx = row[v] > mx ? row[v] : mx;
          if (!isfinite(mx)) mx = 0.0f;
          float r[8];
          float s;
          if (V < 8) {
            s = 0.0f;
            for (int v = 0; v < V; ++v) { row[v] = f32_add(row[v], -mx); s = f32_add(s, expf(row[v])); }
          } else {
            for (int j = 0; j < 8; ++j) { row[j] = f32_add(row[j], -mx); r[j] = expf(row[j]); }
            int i = 8;
            for (; i < V - (V % 8); i += 8)
              for (int j = 0; j < 8; ++j) { row[i + j] = f32_add(row[i + j], -mx); r[j] = f32_add(r[j], expf(row[i + j])); }
            s = f32_add(f32_add(f32_add(r[0], r[1]), f32_add(r[2], r[3])),
                        f32_add(f32_add(r[4], r[5]), f32_add(r[6], r[7])));
            for (; i < V; ++i) { row[i] = f32_add(row[i], -mx); s = f32_add(s, expf(row[i])); }
          }
          const float ls = logf(s);
          for (int v = 0; v < V; ++v) {
            float y = f32_add(row[v], -ls);
            row[v] = y < lo ? lo : (y > 0.0f ? 0.0f : y);
          }
        }
        // argmax (first maximum) and the kept-token list in ascending id
        int am = 0;
        float best = row[0];
        for (int v = 1; v < V; ++v) if (row[v] > best) { best = row[v]; am = v; }
        int nk = 0;
        for (int v = 0; v < V; ++v)
          if (row[v] >= P.token_min_logp || v == am) sm.kept[f][nk++] = (uint8_t)v;
        sm.nkept[f] = (uint8_t)nk;
      }
    }
    CORAL_GSYNC(NT);
  }

  static CORAL_DEV uint32_t rep_beam(Sm& sm, int s) { return sm.sb0[s] != kNone16 ? sm.sb0[s] : sm.sb1[s]; }

  // Sequential log-sum-exp of (logit[b] + p) over up to four member beams in ascending
  // beam index (= the reference's candidate order within one token). Returns min index.
  static CORAL_DEV uint32_t merge_members(Sm& sm, int cur, uint32_t m[4], int n, double p, double& score) {
    // insertion sort of <= 4 indices
    for (int a = 1; a < n; ++a) {
      uint32_t x = m[a];
      int b = a - 1;
      while (b >= 0 && m[b] > x) { m[b + 1] = m[b]; --b; }
      m[b + 1] = x;
    }
    score = d_add(sm.logit[cur][m[0]], p);
    for (int a = 1; a < n; ++a) score = sum_log_scores(score, d_add(sm.logit[cur][m[a]], p));
    return m[0];
  }

  // ---- phase 1: hash the live nodes ----------------------------------------------------
  static CORAL_DEV void build_node_hash(Sm& sm) {
    const int cur = sm.cur;
    CORAL_LANES(NT) {
      for (int i = lane; i < HS; i += NT) { sm.hkey[i] = 0; sm.sb0[i] = kNone16; sm.sb1[i] = kNone16; sm.claimed[i] = 0; }
      if (lane == 0) { sm.nN = 0; sm.n_out = 0; sm.S = 0; sm.S2 = 0; sm.gmax = 0; }
    }
    CORAL_GSYNC(NT);
    const uint32_t nb = sm.nb;
    CORAL_LANES(NT) {
      for (uint32_t b = lane; b < nb; b += NT) {
        const uint32_t mt = sm.meta[cur][b];
        const int s = h_insert(sm, node_key(sm.parent[cur][b], meta_tok(mt)));
        sm.beam_slot[b] = (uint16_t)s;
      }
    }
    CORAL_GSYNC(NT);
    CORAL_LANES(NT) {
      for (uint32_t b = lane; b < nb; b += NT) {
        const int s = sm.beam_slot[b];
        if (meta_lc(sm.meta[cur][b]) == (uint32_t)0xFE) sm.sb0[s] = (uint16_t)b;  // 0xFE = blank marker
        else sm.sb1[s] = (uint16_t)b;
      }
    }
    CORAL_GSYNC(NT);
    CORAL_LANES(NT) {
      for (uint32_t b = lane; b < nb; b += NT) {
        const int s = sm.beam_slot[b];
        const uint32_t a = sm.sb0[s], c = sm.sb1[s];
        const uint32_t first = a < c ? a : c;  // kNone16 is larger than any index
        if (first == b) { const uint32_t j = atom_add(&sm.nN, 1u); sm.ne_slot[j] = (uint16_t)s; }
      }
    }
    CORAL_GSYNC(NT);
  }

  // ---- one frame ------------------------------------------------------------------------
  static CORAL_DEV void frame_step(Sm& sm, const LmView& lm, const DecodeParams& P, const SlotScratch& sc,
                                   const UttIO& io, int f) {
    build_node_hash(sm);
    const int cur = sm.cur;
    const int K = sm.nkept[f];
    const uint32_t nb = sm.nb, nN = sm.nN;
    const uint32_t n_slots = nN * (uint32_t)K + nN;
    OutRec* outs = n_slots <= (uint32_t)OUTC ? sm.outs : sc.outs_g;
    uint16_t* surv = n_slots <= (uint32_t)OUTC ? sm.surv : sc.surv_g;
    if (n_slots > sc.outs_cap && n_slots > (uint32_t)OUTC) { CORAL_LANES(NT) { if (lane == 0) sm.status = -4; } CORAL_GSYNC(NT); return; }
    CORAL_LANES(NT) {
      for (int v = lane; v < P.V; v += NT) sm.krank[v] = 0xFF;
    }
    CORAL_GSYNC(NT);
    CORAL_LANES(NT) {
      for (int k = lane; k < K; k += NT) sm.krank[sm.kept[f][k]] = (uint8_t)k;
      if (lane == 0 && io.stats) { atom_add(&io.stats[0], (unsigned long long)K * nb); atom_add(&io.stats[3], 1ULL); }
    }
    CORAL_GSYNC(NT);

    // -- phase 2a: every (live node, kept token) --------------------------------------
    CORAL_LANES(NT) {
      unsigned long long lmax = 0;
      for (uint32_t i = lane; i < nN * (uint32_t)K; i += NT) {
        const uint32_t j = i / K, k = i % K;
        const int s = sm.ne_slot[j];
        const uint32_t c = sm.kept[f][k];
        const double p = (double)sm.lp[f][c];
        const uint32_t rb = rep_beam(sm, s);
        const uint32_t mt = sm.meta[cur][rb];
        const uint32_t tok_m = meta_tok(mt), wlen_m = meta_wlen(mt), fl_m = meta_flags(mt);
        const uint32_t b0 = sm.sb0[s], b1 = sm.sb1[s];
        OutRec o;
        o.slot = (uint16_t)s;
        o.c = (uint8_t)c;
        o.aux = 0;
        o.child = 0;
        uint32_t mem[4];
        int nm = 0;
        bool valid = true;
        if ((int)c == P.blank_id) {
          if (b0 != kNone16) mem[nm++] = b0;
          if (b1 != kNone16) mem[nm++] = b1;
          const uint32_t first = merge_members(sm, cur, mem, nm, p, o.logit);
          o.order = k * nb + first;
          o.kf = 0;
          o.comb = d_add(o.logit, d_add(sm.lm_raw[cur][rb], partial_score(P, wlen_m, fl_m)));
        } else if ((int)c == P.space_id && wlen_m == 0) {
          valid = false;  // a space after a closed word / at the start never extends
        } else {
          if (b0 != kNone16) mem[nm++] = b0;
          if (b1 != kNone16 && tok_m != c) mem[nm++] = b1;
          const int cs = h_find(sm, node_key(sm.node[cur][rb], c));
          if (cs >= 0) {
            sm.claimed[cs] = 1;
            if (sm.sb1[cs] != kNone16) mem[nm++] = sm.sb1[cs];
            if ((int)c == P.space_id && sm.sb0[cs] != kNone16) mem[nm++] = sm.sb0[cs];
          }
          if (nm == 0) {
            valid = false;
          } else {
            const uint32_t first = merge_members(sm, cur, mem, nm, p, o.logit);
            o.order = k * nb + first;
            if (cs >= 0) {
              const uint32_t crb = rep_beam(sm, cs);
              const uint32_t cmt = sm.meta[cur][crb];
              o.kf = 1;
              o.child = crb;
              o.comb = d_add(o.logit, d_add(sm.lm_raw[cur][crb], partial_score(P, meta_wlen(cmt), meta_flags(cmt))));
            } else if ((int)c == P.space_id) {
              // a word closes: the boundary node is materialised at once (it carries the
              // LM record, like pyctcdecode's cached_lm_scores entry for the new text)
              uint32_t nid = 0, bnd_new = 0;
              double raw_new = sm.lm_raw[cur][rb];
              bool have = false;
              if (trie_find(sc, sm.node[cur][rb], c, nid)) {
                bnd_new = sc.node_info[nid] >> 8;
                if (lm.present) raw_new = sc.bnd[bnd_new].lm_raw;
                have = true;
              }
              if (!have) {
                if (lm.present) {
                  bnd_new = atom_add(&sm.bnd_count, 1u);
                  if (bnd_new >= sc.bnd_cap) { sm.status = -4; bnd_new = 0; }
                  const bool in_lm = (fl_m & kInLm) != 0;
                  const bool oov = (lm.has_unigrams && !(fl_m & kInUni)) || !in_lm;
                  BndRec nr;
                  const double sc_w = lm_word_score(lm, P, sc.bnd[sm.bnd[cur][rb]].st, in_lm ? sm.wid[cur][rb] : 0u,
                                                    oov, false, nr.st, io.stats);
                  nr.lm_raw = d_add(sm.lm_raw[cur][rb], sc_w);
                  raw_new = nr.lm_raw;
                  sc.bnd[bnd_new] = nr;
                }
                bool created;
                nid = trie_get_or_add(sm, sc, sm.node[cur][rb], c, bnd_new, created);
              }
              o.kf = 3;
              o.child = nid;
              o.aux = bnd_new;
              o.comb = d_add(o.logit, d_add(raw_new, 0.0));
            } else {
              // a letter extends the partial word: roll the word hash, probe the lexicon
              uint32_t nfl = fl_m & kDead ? (kDead | kOovPartial) : 0u;
              uint32_t nwid = 0;
              if (lm.present) {
                if (!(fl_m & kDead)) {
                  unsigned long long h = sm.whash[cur][rb];
                  for (int q = 0; q < P.label_ncp[c]; ++q) h = word_hash_push(h, P.label_cps[c][q]);
                  uint32_t lw, lf;
                  if (io.stats) atom_add(&io.stats[4], 1ULL);
                  if (lex_find(lm, h, lw, lf)) {
                    nfl = ((lf & kLexPrefixOfUnigram) ? 0u : kOovPartial) | ((lf & kLexInUnigrams) ? kInUni : 0u) |
                          ((lf & kLexInLm) ? kInLm : 0u);
                    nwid = lw;
                  } else {
                    nfl = kDead | kOovPartial;
                  }
                }
              } else {
                nfl = 0;
              }
              o.kf = (uint8_t)(2u | (nfl << 2));
              o.aux = nwid;
              const double ps = lm.present ? partial_score(P, wlen_m + P.label_ncp[c], nfl) : 0.0;
              o.comb = d_add(o.logit, d_add(sm.lm_raw[cur][rb], ps));
            }
          }
        }
        if (valid) {
          const uint32_t at = atom_add(&sm.n_out, 1u);
          outs[at] = o;
          const unsigned long long ok = ordered_u64(o.comb);
          lmax = ok > lmax ? ok : lmax;
        }
      }
      if (lmax) atom_max_u64(&sm.gmax, lmax);
    }
    CORAL_GSYNC(NT);
    // -- phase 2b: repeats / closing spaces of nodes whose parent is not live -----------
    CORAL_LANES(NT) {
      unsigned long long lmax = 0;
      for (uint32_t j = lane; j < nN; j += NT) {
        const int s = sm.ne_slot[j];
        if (sm.claimed[s]) continue;
        const uint32_t rb = rep_beam(sm, s);
        const uint32_t mt = sm.meta[cur][rb];
        const uint32_t tok_m = meta_tok(mt);
        const uint32_t c = tok_m == kNoTok ? (uint32_t)P.space_id : tok_m;
        const uint32_t k = sm.krank[c];
        if (k == 0xFF) continue;
        uint32_t mem[4];
        int nm = 0;
        const uint32_t b0 = sm.sb0[s], b1 = sm.sb1[s];
        if (b1 != kNone16) mem[nm++] = b1;  // last_char == c (or None/space at the root)
        if ((int)c == P.space_id && b0 != kNone16) mem[nm++] = b0;
        if (nm == 0) continue;
        OutRec o;
        o.slot = (uint16_t)s;
        o.c = (uint8_t)c;
        o.aux = 0;
        o.child = 0;
        o.kf = 0;
        const uint32_t first = merge_members(sm, cur, mem, nm, (double)sm.lp[f][c], o.logit);
        o.order = k * nb + first;
        o.comb = d_add(o.logit, d_add(sm.lm_raw[cur][rb], partial_score(P, meta_wlen(mt), meta_flags(mt))));
        const uint32_t at = atom_add(&sm.n_out, 1u);
        outs[at] = o;
        const unsigned long long ok = ordered_u64(o.comb);
        lmax = ok > lmax ? ok : lmax;
      }
      if (lmax) atom_max_u64(&sm.gmax, lmax);
    }
    CORAL_GSYNC(NT);
    select_and_commit(sm, lm, P, sc, outs, surv, false);
  }

  // ---- prune, trim to beam_width, rank, write the next beam list ------------------------
  static CORAL_DEV void select_and_commit(Sm& sm, const LmView& lm, const DecodeParams& P,
                                          const SlotScratch& sc, OutRec* outs, uint16_t* surv, bool final_pass) {
    const int cur = sm.cur, nxt = cur ^ 1;
    const uint32_t n_out = sm.n_out;
    // max_score + beam_prune_logp, compared in float64
    double maxs;
    {
      unsigned long long u = sm.gmax;
      u = (u >> 63) ? (u & 0x7FFFFFFFFFFFFFFFULL) : ~u;
      union { double d; unsigned long long u; } c;
      c.u = u;
      maxs = c.d;
    }
    const double thr = d_add(maxs, P.beam_prune_logp);
    CORAL_LANES(NT) {
      for (uint32_t i = lane; i < n_out; i += NT)
        if (outs[i].comb >= thr) surv[atom_add(&sm.S, 1u)] = (uint16_t)i;
    }
    CORAL_GSYNC(NT);
    uint32_t S = sm.S;
    if (S > (uint32_t)P.beam_width) {
      // radix select of the beam_width best by (score desc, first-candidate index asc)
      CORAL_LANES(NT) { if (lane == 0) { sm.sel_prefix = 0; sm.sel_mask = 0; sm.sel_need = (uint32_t)P.beam_width; } }
      CORAL_GSYNC(NT);
      for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        CORAL_LANES(NT) { for (int i = lane; i < 256; i += NT) sm.hist[i] = 0; }
        CORAL_GSYNC(NT);
        const unsigned long long pre = sm.sel_prefix, msk = sm.sel_mask;
        CORAL_LANES(NT) {
          for (uint32_t i = lane; i < S; i += NT) {
            const unsigned long long k = ordered_u64(outs[surv[i]].comb);
            if ((k & msk) == pre) atom_add(&sm.hist[(k >> shift) & 255], 1u);
          }
        }
        CORAL_GSYNC(NT);
        CORAL_LANES(NT) {
          if (lane == 0) {
            uint32_t cum = 0, need = sm.sel_need;
            int d = 255;
            for (; d > 0; --d) { if (cum + sm.hist[d] >= need) break; cum += sm.hist[d]; }
            sm.sel_need = need - cum;
            sm.sel_eq = sm.hist[d];
            sm.sel_prefix = pre | ((unsigned long long)d << shift);
            sm.sel_mask = msk | (255ULL << shift);
          }
        }
        CORAL_GSYNC(NT);
      }
      const unsigned long long kth = sm.sel_prefix;
      // ties at the cut: keep the sel_need smallest candidate indices
      uint32_t ord_cut = 0xFFFFFFFFu;
      if (sm.sel_eq != sm.sel_need) {
        CORAL_LANES(NT) {
          if (lane == 0) {
            // rare: select by repeated minimum
            uint32_t last = 0; bool first = true; uint32_t cut = 0;
            for (uint32_t r = 0; r < sm.sel_need; ++r) {
              uint32_t best = 0xFFFFFFFFu;
              for (uint32_t i = 0; i < S; ++i) {
                const OutRec& o = outs[surv[i]];
                if (ordered_u64(o.comb) == kth && (first || o.order > last) && o.order < best) best = o.order;
              }
              last = best; first = false; cut = best;
            }
            sm.sel_need = cut;  // reuse as the order cut
          }
        }
        CORAL_GSYNC(NT);
        ord_cut = sm.sel_need;
      }
      // losers are marked kNone16 in place; the ranking below skips them
      CORAL_LANES(NT) {
        for (uint32_t i = lane; i < S; i += NT) {
          const OutRec& o = outs[surv[i]];
          const unsigned long long k = ordered_u64(o.comb);
          if (!(k > kth || (k == kth && o.order <= ord_cut))) surv[i] = (uint16_t)kNone16;
        }
      }
      CORAL_GSYNC(NT);
    }
    // rank every survivor and write it at its position in the other beam buffer
    CORAL_LANES(NT) {
      for (uint32_t i = lane; i < S; i += NT) {
        const uint32_t oi = surv[i];
        if (oi == kNone16) continue;
        const OutRec o = outs[oi];
        uint32_t r = 0;
        for (uint32_t j = 0; j < S; ++j) {
          const uint32_t oj = surv[j];
          if (oj == kNone16) continue;
          const double cj = outs[oj].comb;
          r += (cj > o.comb) || (cj == o.comb && outs[oj].order < o.order);
        }
        if (final_pass) {
          sm.logit[nxt][r] = o.logit;
          sm.lm_raw[nxt][r] = o.comb;   // final: combined score travels in lm_raw
          sm.node[nxt][r] = o.child;    // final: text node
          continue;
        }
        const int s = o.slot;
        const uint32_t rb = rep_beam(sm, s);
        const uint32_t kind = o.kf & 3u;
        const uint32_t c = o.c;
        const uint32_t lc = (int)c == P.blank_id ? 0xFEu : c;
        if (kind == 0 || kind == 1) {
          const uint32_t src = kind == 0 ? rb : o.child;
          const uint32_t mt = sm.meta[cur][src];
          sm.node[nxt][r] = sm.node[cur][src];
          sm.parent[nxt][r] = sm.parent[cur][src];
          sm.bnd[nxt][r] = sm.bnd[cur][src];
          sm.wid[nxt][r] = sm.wid[cur][src];
          sm.whash[nxt][r] = sm.whash[cur][src];
          sm.lm_raw[nxt][r] = sm.lm_raw[cur][src];
          sm.meta[nxt][r] = meta_pack(meta_tok(mt), lc, meta_flags(mt), meta_wlen(mt));
        } else if (kind == 2) {
          bool created;
          const uint32_t nid = trie_get_or_add(sm, sc, sm.node[cur][rb], c, 0u, created);
          unsigned long long h = sm.whash[cur][rb];
          for (int q = 0; q < P.label_ncp[c]; ++q) h = word_hash_push(h, P.label_cps[c][q]);
          sm.node[nxt][r] = nid;
          sm.parent[nxt][r] = sm.node[cur][rb];
          sm.bnd[nxt][r] = sm.bnd[cur][rb];
          sm.wid[nxt][r] = o.aux;
          sm.whash[nxt][r] = h;
          sm.lm_raw[nxt][r] = sm.lm_raw[cur][rb];
          sm.meta[nxt][r] = meta_pack(c, lc, (uint32_t)(o.kf >> 2), meta_wlen(sm.meta[cur][rb]) + P.label_ncp[c]);
        } else {
          sm.node[nxt][r] = o.child;
          sm.parent[nxt][r] = sm.node[cur][rb];
          sm.bnd[nxt][r] = o.aux;
          sm.wid[nxt][r] = 0;
          sm.whash[nxt][r] = kWordHashSeed;
          sm.lm_raw[nxt][r] = lm.present ? sc.bnd[o.aux].lm_raw : 0.0;
          sm.meta[nxt][r] = meta_pack(c, lc, 0u, 0u);
        }
        sm.logit[nxt][r] = o.logit;
      }
    }
    CORAL_GSYNC(NT);
    CORAL_LANES(NT) {
      if (lane == 0) {
        uint32_t cnt = S > (uint32_t)P.beam_width ? (uint32_t)P.beam_width : S;
        sm.nb = cnt;
        sm.cur = nxt;
      }
    }
    CORAL_GSYNC(NT);
  }

  // ---- end of utterance ------------------------------------------------
  static CORAL_DEV void finalize(Sm& sm, const LmView& lm, const DecodeParams& P, const SlotScratch& sc,
                                 const UttIO& io) {
    build_node_hash(sm);
    const int cur = sm.cur;
    const uint32_t nN = sm.nN;
    OutRec* outs = nN <= (uint32_t)OUTC ? sm.outs : sc.outs_g;
    uint16_t* surv = nN <= (uint32_t)OUTC ? sm.surv : sc.surv_g;
    // pass A: nodes with an open word (or the root) lead; they absorb their closed-word child
    for (int pass = 0; pass < 2; ++pass) {
      CORAL_LANES(NT) {
        unsigned long long lmax = 0;
        for (uint32_t j = lane; j < nN; j += NT) {
          const int s = sm.ne_slot[j];
          const uint32_t rb = rep_beam(sm, s);
          const uint32_t mt = sm.meta[cur][rb];
          const bool open_or_root = meta_wlen(mt) > 0 || meta_tok(mt) == kNoTok;
          if (pass == 0 ? !open_or_root : (open_or_root || sm.claimed[s])) continue;
          uint32_t mem[4];
          int nm = 0;
          if (sm.sb0[s] != kNone16) mem[nm++] = sm.sb0[s];
          if (sm.sb1[s] != kNone16) mem[nm++] = sm.sb1[s];
          if (pass == 0 && meta_wlen(mt) > 0) {
            const int cs = h_find(sm, node_key(sm.node[cur][rb], (uint32_t)P.space_id));
            if (cs >= 0) {
              sm.claimed[cs] = 1;
              if (sm.sb0[cs] != kNone16) mem[nm++] = sm.sb0[cs];
              if (sm.sb1[cs] != kNone16) mem[nm++] = sm.sb1[cs];
            }
          }
          OutRec o;
          const uint32_t first = merge_members(sm, cur, mem, nm, 0.0, o.logit);
          // logit + 0.0 is exact, so merge_members' "+ p" leaves the scores untouched
          o.order = first;
          const uint32_t last = mem[nm - 1];  // the later candidate's tuple is the one stored
          const uint32_t lmt = sm.meta[cur][last];
          double comb = o.logit;
          if (lm.present) {
            const bool open = meta_wlen(lmt) > 0;
            const uint32_t lfl = meta_flags(lmt);
            const bool in_lm = open && (lfl & kInLm);
            const bool oov = !open || (lm.has_unigrams && !(lfl & kInUni)) || !in_lm;
            LmState out;
            const double sw = lm_word_score(lm, P, sc.bnd[sm.bnd[cur][last]].st, in_lm ? sm.wid[cur][last] : 0u,
                                            oov, true, out, io.stats);
            comb = d_add(o.logit, d_add(d_add(sm.lm_raw[cur][last], sw), 0.0));
          }
          o.comb = comb;
          // text node: the open-word node itself, or the parent of a closed-word node
          o.child = open_or_root ? sm.node[cur][rb] : sm.parent[cur][rb];
          o.slot = (uint16_t)s;
          o.c = 0;
          o.kf = 0;
          o.aux = 0;
          outs[atom_add(&sm.n_out, 1u)] = o;
          const unsigned long long ok = ordered_u64(o.comb);
          lmax = ok > lmax ? ok : lmax;
        }
        if (lmax) atom_max_u64(&sm.gmax, lmax);
      }
      CORAL_GSYNC(NT);
    }
    select_and_commit(sm, lm, P, sc, outs, surv, true);
    const int fin = sm.cur;
    const uint32_t nf = sm.nb;
    CORAL_LANES(NT) {
      if (lane == 0) { *io.out_n = (int32_t)nf; *io.out_status = sm.status; }
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
      }
    }
    CORAL_GSYNC(NT);
  }

  // ---- whole utterance ---------------------------------------------------------------------
  static CORAL_DEV void decode(Sm& sm, const LmView& lm, const DecodeParams& P, SlotScratch& sc, const UttIO& io) {
    CORAL_LANES(NT) {
      if (lane == 0) {
        sm.cur = 0;
        sm.nb = 1;
        sm.status = 0;
        sm.node_count = 1;
        sm.bnd_count = 1;
        sm.logit[0][0] = 0.0;
        sm.lm_raw[0][0] = 0.0;
        sm.whash[0][0] = kWordHashSeed;
        sm.node[0][0] = 0;
        sm.parent[0][0] = kNoNode;
        sm.bnd[0][0] = 0;
        sm.wid[0][0] = 0;
        sm.meta[0][0] = meta_pack(kNoTok, 0xFFu, 0u, 0u);
        sc.node_parent[0] = kNoNode;
        sc.node_info[0] = kNoTok;
        if (lm.present) {
          BndRec r0;
          r0.lm_raw = 0.0;
          if (P.score_boundary) lm_begin_sentence(lm, r0.st); else lm_null_context(r0.st);
          sc.bnd[0] = r0;
        }
      }
    }
    CORAL_GSYNC(NT);
    for (int t0 = 0; t0 < io.T; t0 += kChunk) {
      const int nf = io.T - t0 < kChunk ? io.T - t0 : kChunk;
      stage_frames(sm, P, io, t0, nf);
      for (int f = 0; f < nf; ++f) {
        frame_step(sm, lm, P, sc, io, f);
        if (sm.status != 0) break;
      }
      if (sm.status != 0) break;
    }
    if (sm.status != 0) {
      CORAL_LANES(NT) { if (lane == 0) { *io.out_n = 0; *io.out_status = sm.status; } }
      CORAL_GSYNC(NT);
      return;
    }
    finalize(sm, lm, P, sc, io);
  }
};

}  // namespace coral
