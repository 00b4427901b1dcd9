// C-ABI: decoder handle and the batched prefix-beam-search launch.
// The per-utterance algorithm lives in beam_core.h; this file owns the persistent
// kernel (one thread group per utterance at a time, work-stealing over the batch),
// the HBM scratch arenas and pyctcdecode's probabilities-vs-logits detection.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "handles.h"

#include "beam_launch.cuh"

using namespace coral;

// threads per utterance of the two wide-beam instantiations (tuning: profiles/r2_beam_kernel_tuning.md)
#ifndef CORAL_NT_BEAM256
#define CORAL_NT_BEAM256 256
#endif
#ifndef CORAL_NT_BEAM512
#define CORAL_NT_BEAM512 512
#endif
#ifndef CORAL_OUTC_BEAM128
#define CORAL_OUTC_BEAM128 320
#endif
#ifndef CORAL_OUTC_BEAM256
#define CORAL_OUTC_BEAM256 416
#endif
#ifndef CORAL_OUTC_BEAM512
#define CORAL_OUTC_BEAM512 1024
#endif

extern "C" {

int32_t coral_decoder_create(const uint32_t* label_cps, const int32_t* label_offsets, int32_t n_labels,
                             int32_t blank_id, int32_t space_id, const coral_lm* lm, const uint32_t* unigram_cps,
                             const int64_t* unigram_offsets, int64_t n_unigrams, int32_t device,
                             coral_decoder** out) {
  if (!out || !label_offsets || (!label_cps && n_labels > 0)) return fail(CORAL_EARG, "coral_decoder_create: null argument");
  *out = nullptr;
  if (n_labels < 1 || n_labels > kVMax)
    return fail(CORAL_EARG, "alphabet size must be in [1, 64] in this build");
  if (blank_id < 0 || blank_id >= n_labels) return fail(CORAL_EARG, "blank_id outside the alphabet");
  if (space_id >= n_labels) return fail(CORAL_EARG, "space_id outside the alphabet");
  if (lm && lm->device != device) return fail(CORAL_EARG, "LM and decoder must live on the same device");
  coral_decoder* d = new coral_decoder();
  memset(&d->P, 0, sizeof(d->P));
  d->lm = lm;
  d->device = device;
  d->P.V = n_labels;
  d->P.blank_id = blank_id;
  d->P.space_id = space_id;
  d->P.alpha = 0.5;
  d->P.beta = 1.5;
  d->P.unk_score_offset = -10.0;
  d->P.score_boundary = 1;
  d->P.log_base_change = 2.302585092994046;  // 1.0 / np.log10(np.e), UP:pyctcdecode constants.py
  for (int v = 0; v < n_labels; ++v) {
    const int n = label_offsets[v + 1] - label_offsets[v];
    if (n < 0 || n > kMaxLabelCps) { delete d; return fail(CORAL_EARG, "alphabet label longer than 8 code points"); }
    if ((n == 0) != (v == blank_id)) { delete d; return fail(CORAL_EARG, "exactly the blank label must be empty"); }
    d->P.label_ncp[v] = (uint8_t)n;
    for (int q = 0; q < n; ++q) d->P.label_cps[v][q] = label_cps[label_offsets[v] + q];
  }
  if (lm) {
    std::vector<std::u32string> uni;
    if (n_unigrams >= 0) {
      uni.reserve((size_t)n_unigrams);
      for (int64_t i = 0; i < n_unigrams; ++i)
        uni.emplace_back(reinterpret_cast<const char32_t*>(unigram_cps + unigram_offsets[i]),
                         (size_t)(unigram_offsets[i + 1] - unigram_offsets[i]));
    }
    std::string err;
    std::vector<std::u32string> labels;
    for (int v = 0; v < n_labels; ++v)
      labels.emplace_back(reinterpret_cast<const char32_t*>(label_cps + label_offsets[v]),
                          (size_t)(label_offsets[v + 1] - label_offsets[v]));
    const int rc = build_lexicon(lm->host, n_unigrams >= 0 ? &uni : nullptr, d->lex, err, &labels);
    if (rc != 0) { delete d; return fail(rc, err); }
    DeviceGuard g(device);
    const size_t lb = d->lex.lex.size() * sizeof(LexSlot);
    const size_t ob = d->lex.child_ok.size() * sizeof(uint64_t);
    cudaError_t e;
    if ((e = cudaMalloc(&d->d_lex, lb)) != cudaSuccess ||
        (e = cudaMemcpy(d->d_lex, d->lex.lex.data(), lb, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (ob && ((e = cudaMalloc(&d->d_lex_ok, ob)) != cudaSuccess ||
                (e = cudaMemcpy(d->d_lex_ok, d->lex.child_ok.data(), ob, cudaMemcpyHostToDevice)) != cudaSuccess))) {
      std::string m = std::string("uploading lexicon: ") + cudaGetErrorString(e);
      coral_decoder_free(d);
      return fail(CORAL_ECUDA, m);
    }
    d->device_bytes = lb + ob;
  }
  *out = d;
  return CORAL_OK;
}

int32_t coral_decoder_free(coral_decoder* d) {
  if (!d) return CORAL_OK;
  DeviceGuard g(d->device);
  cudaDeviceSynchronize();
  if (d->d_lex) cudaFree(d->d_lex);
  if (d->d_lex_ok) cudaFree(d->d_lex_ok);
  for (auto& kv : d->scratch) {
    if (kv.second.d_scratch) cudaFree(kv.second.d_scratch);
    if (kv.second.d_work) cudaFree(kv.second.d_work);
  }
  delete d;
  return CORAL_OK;
}

int32_t coral_decoder_set_params(coral_decoder* d, double alpha, double beta, double unk_score_offset,
                                 int32_t score_boundary) {
  if (!d) return fail(CORAL_EARG, "coral_decoder_set_params: null handle");
  d->P.alpha = alpha;
  d->P.beta = beta;
  d->P.unk_score_offset = unk_score_offset;
  d->P.score_boundary = score_boundary ? 1 : 0;
  return CORAL_OK;
}

int32_t coral_decoder_info(const coral_decoder* d, uint64_t* lexicon_entries, uint64_t* device_bytes) {
  if (!d) return fail(CORAL_EARG, "coral_decoder_info: null handle");
  if (lexicon_entries) *lexicon_entries = d->lex.n_entries;
  if (device_bytes) {
    uint64_t sb = 0;
    for (const auto& kv : d->scratch) sb += kv.second.scratch_bytes;
    *device_bytes = d->device_bytes + sb;
  }
  return CORAL_OK;
}

int32_t coral_ctc_beam_decode(coral_decoder* dec, const float* logits_dev, const int32_t* lengths_dev,
                              const int32_t* order_dev, const int64_t* frame_offsets_dev, int32_t B, int32_t T_max,
                              int32_t V, int32_t beam_width,
                              double beam_prune_logp, double token_min_logp, int32_t prune_history,
                              int32_t input_mode, int32_t n_best, int32_t* out_n_beams_dev,
                              double* out_logit_score_dev, double* out_lm_score_dev, uint8_t* out_tokens_dev,
                              int32_t* out_lens_dev, int32_t* out_status_dev, uint64_t* stats_dev,
                              const int32_t* ready_dev, int32_t ready_chunk, int32_t* out_word_frames_dev,
                              int32_t* out_word_counts_dev, int32_t max_words, void* stream) {
  if (!dec) return fail(CORAL_EARG, "coral_ctc_beam_decode: null decoder");
  if ((out_word_frames_dev != nullptr) != (out_word_counts_dev != nullptr) || (out_word_frames_dev && max_words < 1))
    return fail(CORAL_EARG, "word frames need both output buffers and max_words >= 1");
  if (ready_dev && ready_chunk < 1) return fail(CORAL_EARG, "ready_chunk must be >= 1 with a ready counter");
  if (B < 0 || T_max < 0) return fail(CORAL_EARG, "negative batch or frame count");
  if (V != dec->P.V)
    return fail(CORAL_EARG, "Input logits have vocabulary size " + std::to_string(V) + ", but the alphabet is size " +
                                std::to_string(dec->P.V) + ". Need logits of shape: (time, vocabulary)");
  if (beam_width < 1 || beam_width > 512) return fail(CORAL_EARG, "beam_width must be in [1, 512]");
  if (n_best < 1 || n_best > beam_width) return fail(CORAL_EARG, "n_best must be in [1, beam_width]");
  if (input_mode < 0 || input_mode > 2) return fail(CORAL_EARG, "input_mode must be 0, 1 or 2");
  if (B == 0) return CORAL_OK;
  if (!logits_dev || !lengths_dev || !out_n_beams_dev || !out_logit_score_dev || !out_lm_score_dev ||
      !out_tokens_dev || !out_lens_dev || !out_status_dev)
    return fail(CORAL_EARG, "coral_ctc_beam_decode: null buffer");
  DeviceGuard g(dec->device);
  cudaStream_t st = (cudaStream_t)stream;

  BeamLaunch L;
  memset(&L, 0, sizeof(L));
  L.P = dec->P;
  L.P.beam_width = beam_width;
  L.P.n_best = n_best;
  L.P.T_max = T_max > 0 ? T_max : 1;
  L.P.input_mode = input_mode;
  // pyctcdecode _prune_history: min_n_history = max(1, lm_order - 1), lm_order = 1 without a language model
  L.P.prune_history = prune_history ? std::max(1, (dec->lm ? dec->lm->host.order : 1) - 1) : 0;
  {
    // pinned host logits are read in place (zero-copy): plain loads + L2 prefetch of the next
    // frames (mode 2, default: measured 16.3 ms per 8192 utterances against 15.3 ms for
    // HBM-resident logits); CORAL_HOST_INPUT_MODE=1 selects uncached loads (46 ms) -- a knob
    // kept for diagnosing stale-line suspicions, see DESIGN.md
    cudaPointerAttributes pa;
    L.P.host_input = 0;
    if (cudaPointerGetAttributes(&pa, logits_dev) == cudaSuccess && pa.type == cudaMemoryTypeHost) {
      L.P.host_input = 2;
      if (const char* e = getenv("CORAL_HOST_INPUT_MODE")) { const int v = atoi(e); if (v == 1 || v == 2) L.P.host_input = v; }
    } else {
      cudaGetLastError();  // an unregistered host pointer sets a sticky-free error: clear it
    }
  }
  L.P.token_min_logp = (float)token_min_logp;
  L.P.beam_prune_logp = beam_prune_logp;
  set_bucket_scale(L.P);
  if (dec->lm) {
    L.lm = make_view(dec->lm->host, dec->lex, dec->lm->d_uni, dec->lm->d_ng, dec->d_lex, dec->d_lex_ok);
  } else {
    L.lm.present = 0;
  }
  L.logits = logits_dev;
  L.lengths = lengths_dev;
  L.order = order_dev;
  L.frame_off = frame_offsets_dev;
  L.B = B;
  L.out_n = out_n_beams_dev;
  L.out_logit = out_logit_score_dev;
  L.out_comb = out_lm_score_dev;
  L.out_tokens = out_tokens_dev;
  L.out_len = out_lens_dev;
  L.out_status = out_status_dev;
  L.stats = reinterpret_cast<unsigned long long*>(stats_dev);
  L.ready = ready_dev;
  L.ready_chunk = ready_chunk > 0 ? ready_chunk : 1;
  L.ready_timeout = 40LL << 30;  // about 20 s; CORAL_READY_TIMEOUT_CYCLES overrides (tests)
  if (const char* e = getenv("CORAL_READY_TIMEOUT_CYCLES")) { const long long v = atoll(e); if (v > 0) L.ready_timeout = v; }
  L.out_frames = out_word_frames_dev;
  L.out_nwords = out_word_counts_dev;
  L.max_words = max_words;

  if (out_word_frames_dev)  // the instantiations that also track pyctcdecode's word frames (beam_frames.cu)
    return launch_beam_frames(dec, L, B, beam_width, st);

  if (beam_width <= 32) return launch_beam<32, 32, 128>(dec, L, B, st);
  if (beam_width <= 64) return launch_beam<64, 64, 192>(dec, L, B, st);
  // up to 104 beams (pyctcdecode's default is 100): arrays sized so that EIGHT thread groups fit
  // on an SM (28 KB of shared memory, 64 registers) against the 128-beam instantiation's six.
  // Resident groups per SM 4 / 5 / 6 / 7 / 8: 22.4 / 19.0 / 17.1 / 15.5 / 15.1 ms per 8192 utterances
  if (beam_width <= 104) return launch_beam<128, 104, 208>(dec, L, B, st);
  if (beam_width <= 128) return launch_beam<128, 128, CORAL_OUTC_BEAM128>(dec, L, B, st);
  if (beam_width <= 256) return launch_beam<CORAL_NT_BEAM256, 256, CORAL_OUTC_BEAM256>(dec, L, B, st);
  return launch_beam<CORAL_NT_BEAM512, 512, CORAL_OUTC_BEAM512>(dec, L, B, st);
}

}  // extern "C"
