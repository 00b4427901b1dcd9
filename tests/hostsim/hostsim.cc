// TEST-ONLY host simulation of the beam-search kernel logic.
//
// Compiles coral_b200/csrc/beam_core.h with -DCORAL_HOSTSIM so that the lane loop of each
// phase runs sequentially on the CPU. It lets `pytest -m "not gpu"` check the device
// algorithm (node trie, gather-merge, LM records, selection) against the oracle without a
// GPU. It is built into tests/hostsim/_build/ and loaded only by tests/; nothing under
// coral_b200/ references it, and it is not a fallback for anything.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../coral_b200/csrc/beam_core.h"
#include "../../coral_b200/csrc/lm_host.h"

using namespace coral;

struct HsDecoder {
  HostLm lm;
  HostLexicon lx;
  bool has_lm = false;
  DecodeParams P;
  std::string err;
};

static thread_local std::string g_err;

extern "C" {

const char* hs_last_error() { return g_err.c_str(); }

void* hs_create(const char* arpa_path, const uint32_t* label_cps, const int32_t* label_off, int n_labels,
                int blank_id, int space_id, const uint32_t* uni_cps, const int64_t* uni_off, int64_t n_uni) {
  HsDecoder* d = new HsDecoder();
  memset(&d->P, 0, sizeof(d->P));
  if (n_labels > kVMax) { g_err = "alphabet too large"; delete d; return nullptr; }
  d->P.V = n_labels;
  d->P.blank_id = blank_id;
  d->P.space_id = space_id;
  for (int v = 0; v < n_labels; ++v) {
    int n = label_off[v + 1] - label_off[v];
    if (n > kMaxLabelCps) { g_err = "label too long"; delete d; return nullptr; }
    d->P.label_ncp[v] = (uint8_t)n;
    for (int q = 0; q < n; ++q) d->P.label_cps[v][q] = label_cps[label_off[v] + q];
  }
  if (arpa_path) {
    const int rc = is_kenlm_binary(arpa_path) ? load_kenlm_binary(arpa_path, d->lm, g_err) : load_arpa(arpa_path, d->lm, g_err);
    if (rc != 0) { delete d; return nullptr; }
    std::vector<std::u32string> uni;
    if (n_uni >= 0) {
      for (int64_t i = 0; i < n_uni; ++i)
        uni.emplace_back((const char32_t*)(uni_cps + uni_off[i]), (size_t)(uni_off[i + 1] - uni_off[i]));
    }
    std::vector<std::u32string> labels;
    for (int v = 0; v < n_labels; ++v)
      labels.emplace_back((const char32_t*)(label_cps + label_off[v]), (size_t)(label_off[v + 1] - label_off[v]));
    if (build_lexicon(d->lm, n_uni >= 0 ? &uni : nullptr, d->lx, g_err, &labels) != 0) { delete d; return nullptr; }
    d->has_lm = true;
  }
  return d;
}

void hs_free(void* h) { delete (HsDecoder*)h; }

int hs_lm_info(void* h, int* order, uint64_t* counts, uint64_t* n_vocab, uint64_t* n_lex) {
  HsDecoder* d = (HsDecoder*)h;
  *order = d->lm.order;
  for (int i = 0; i < d->lm.order; ++i) counts[i] = d->lm.loaded[i];
  *n_vocab = d->lm.uni.size();
  *n_lex = d->lx.n_entries;
  return 0;
}

// score a sentence of words (utf-32) like kenlm.Model.full_scores: per-word log10 prob
int hs_score_sentence(void* h, const uint32_t* cps, const int64_t* word_off, int n_words, int bos, int eos,
                      float* out_prob, int32_t* out_oov) {
  HsDecoder* d = (HsDecoder*)h;
  LmView lm = make_view(d->lm, d->lx, d->lm.uni.data(), d->lm.ng.data(), d->lx.lex.data());
  LmState st, nx;
  if (bos) lm_begin_sentence(lm, st); else lm_null_context(st);
  for (int i = 0; i < n_words; ++i) {
    uint64_t hsh = kWordHashSeed;
    for (int64_t q = word_off[i]; q < word_off[i + 1]; ++q) hsh = word_hash_push(hsh, cps[q]);
    uint32_t wid = 0, fl = 0;
    bool in_lm = lex_find(lm, hsh, wid, fl) && (fl & kLexInLm);
    if (!in_lm) wid = 0;
    out_prob[i] = lm_base_score(lm, st, wid, nx);
    out_oov[i] = in_lm ? 0 : 1;
    st = nx;
  }
  if (eos) out_prob[n_words] = lm_base_score(lm, st, lm.eos_id, nx);
  return 0;
}

// numpy-order float32 pairwise sum (the kernel's input classification), for a direct check
float hs_pairwise_sum(const float* a, int n) { return np_pairwise_sum(a, n); }

}  // extern "C"

struct FrameOut {
  int32_t* frames;
  int32_t* nwords;
  int32_t max_words;
};

template <int NT, int BW, int OUTC, bool FRAMES>
static int run_t(HsDecoder* d, const DecodeParams& P, const float* logits, int T, int is_prob, int32_t* out_n,
                 double* out_logit, double* out_comb, uint8_t* out_tokens, int32_t* out_len,
                 unsigned long long* stats, int n_utt_repeat, FrameOut fo) {
  using Dec = BeamDecoder<NT, BW, OUTC, FRAMES>;
  typename Dec::Sm* sm = new typename Dec::Sm();
  LmView lm;
  memset(&lm, 0, sizeof(lm));
  if (d->has_lm) lm = make_view(d->lm, d->lx, d->lm.uni.data(), d->lm.ng.data(), d->lx.lex.data(), d->lx.child_ok.data());
  SlotScratch sc;
  sc.node_cap = (uint32_t)((size_t)P.beam_width * (size_t)(T > 0 ? T : 1) + 64);
  sc.bnd_cap = (uint32_t)((size_t)P.beam_width * (size_t)(T > 0 ? T : 1) + 64);
  sc.outs_cap = (uint32_t)(P.beam_width * (P.V + 1) + 64);
  std::vector<uint32_t> node_parent(sc.node_cap), node_info(sc.node_cap);
  std::vector<BndRec> bnd(sc.bnd_cap);
  sc.outs_cap = (sc.outs_cap + 3) & ~3u;
  std::vector<unsigned long long> g_key(sc.outs_cap);
  std::vector<double> g_logit(sc.outs_cap);
  // order, aux, child, info and -- addressed from info, see SlotScratch -- hv_sorted, hv_masks, history records
  std::vector<uint32_t> g_block((size_t)sc.outs_cap * 5 + SlotScratch::kHvMaskBytes / 4 + ((sc.outs_cap + 15) & ~15u) / 4 + 4 +
                                (P.prune_history ? (size_t)sc.bnd_cap * sizeof(HistRec) / 4 : 0));
  uint32_t* g_order = g_block.data();
  uint32_t* g_aux = g_order + sc.outs_cap;
  uint32_t* g_child = g_aux + sc.outs_cap;
  uint32_t* g_info = g_child + sc.outs_cap;
  std::vector<FrameRec> wf(FRAMES ? sc.node_cap : 1);
  sc.wf = wf.data();
  sc.wf_cap = (uint32_t)wf.size();
  sc.node_parent = node_parent.data();
  sc.node_info = node_info.data();
  std::vector<float> rowsum((size_t)(T > 0 ? T : 1) * 9 + 16);
  sc.rowsum = rowsum.data();
  sc.bnd = bnd.data();
  sc.outs_g.key = g_key.data();
  sc.outs_g.logit = g_logit.data();
  sc.outs_g.order = g_order;
  sc.outs_g.aux = g_aux;
  sc.outs_g.child = g_child;
  sc.outs_g.info = g_info;
  int32_t status = 0;
  // decode the same utterance n_utt_repeat times on the same slot: exercises slot reuse
  for (int rep = 0; rep < n_utt_repeat; ++rep) {
    UttIO io;
    io.logits = logits;
    io.T = T;
    io.out_n = out_n;
    io.out_logit = out_logit;
    io.out_comb = out_comb;
    io.out_tokens = out_tokens;
    io.out_len = out_len;
    io.out_status = &status;
    io.out_frames = fo.frames;
    io.out_nwords = fo.nwords;
    io.max_words = fo.max_words;
    io.stats = rep == 0 ? stats : nullptr;
    Dec::decode(*sm, lm, P, sc, io);
  }
  delete sm;
  return status;
}

template <int NT, int BW, int OUTC>
static int run(HsDecoder* d, const DecodeParams& P, const float* logits, int T, int is_prob, int32_t* out_n,
               double* out_logit, double* out_comb, uint8_t* out_tokens, int32_t* out_len,
               unsigned long long* stats, int n_utt_repeat, FrameOut fo) {
  if (fo.frames) return run_t<NT, BW, OUTC, true>(d, P, logits, T, is_prob, out_n, out_logit, out_comb, out_tokens, out_len, stats, n_utt_repeat, fo);
  return run_t<NT, BW, OUTC, false>(d, P, logits, T, is_prob, out_n, out_logit, out_comb, out_tokens, out_len, stats, n_utt_repeat, fo);
}

extern "C" {

int hs_decode(void* h, const float* logits, int T, int is_prob, int beam_width, double beam_prune_logp,
              double token_min_logp, double alpha, double beta, double unk_score_offset, int score_boundary,
              double log_base_change, int input_mode, int n_best, int variant, int repeat, int32_t* out_n,
              double* out_logit, double* out_comb, uint8_t* out_tokens, int32_t* out_len,
              unsigned long long* stats, int32_t* out_frames, int32_t* out_nwords, int32_t max_words) {
  HsDecoder* d = (HsDecoder*)h;
  const FrameOut fo{out_frames, out_nwords, max_words};
  DecodeParams P = d->P;
  P.beam_width = beam_width;
  P.n_best = n_best;
  P.T_max = T > 0 ? T : 1;
  P.input_mode = input_mode;
  P.score_boundary = score_boundary;
  P.token_min_logp = (float)token_min_logp;
  P.beam_prune_logp = beam_prune_logp;
  {
    // variant >= 100 selects prune_history=True on variant - 100 (keeps the C signature stable)
    const bool ph = variant >= 100;
    if (ph) variant -= 100;
    P.prune_history = ph ? (d->has_lm ? (d->lm.order - 1 > 1 ? d->lm.order - 1 : 1) : 1) : 0;
  }
  P.alpha = alpha;
  P.beta = beta;
  P.unk_score_offset = unk_score_offset;
  P.log_base_change = log_base_change;
  set_bucket_scale(P);
  switch (variant) {
    case 0: if (beam_width > 128) break; return run<32, 128, 256>(d, P, logits, T, is_prob, out_n, out_logit, out_comb, out_tokens, out_len, stats, repeat, fo);
    case 1: if (beam_width > 32) break; return run<32, 32, 128>(d, P, logits, T, is_prob, out_n, out_logit, out_comb, out_tokens, out_len, stats, repeat, fo);
    case 2: if (beam_width > 512) break; return run<128, 512, 1024>(d, P, logits, T, is_prob, out_n, out_logit, out_comb, out_tokens, out_len, stats, repeat, fo);
    case 3: if (beam_width > 128) break; return run<64, 128, 128>(d, P, logits, T, is_prob, out_n, out_logit, out_comb, out_tokens, out_len, stats, repeat, fo);
    case 4: if (beam_width > 104) break; return run<128, 104, 208>(d, P, logits, T, is_prob, out_n, out_logit, out_comb, out_tokens, out_len, stats, repeat, fo);  // the production default
    // the other production instantiations (coral_b200/csrc/beam.cu: coral_ctc_beam_decode)
    case 5: if (beam_width > 256) break; return run<256, 256, 640>(d, P, logits, T, is_prob, out_n, out_logit, out_comb, out_tokens, out_len, stats, repeat, fo);
    case 6: if (beam_width > 512) break; return run<256, 512, 1280>(d, P, logits, T, is_prob, out_n, out_logit, out_comb, out_tokens, out_len, stats, repeat, fo);
    case 7: if (beam_width > 128) break; return run<128, 128, 320>(d, P, logits, T, is_prob, out_n, out_logit, out_comb, out_tokens, out_len, stats, repeat, fo);
    case 8: if (beam_width > 64) break; return run<64, 64, 192>(d, P, logits, T, is_prob, out_n, out_logit, out_comb, out_tokens, out_len, stats, repeat, fo);
    default: break;
  }
  g_err = "unsupported variant / beam width";
  return -1;
}

}  // extern "C"
