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

namespace coral {

struct BeamLaunch {
  DecodeParams P;
  LmView lm;
  const float* logits;
  const int32_t* lengths;
  const int32_t* order;
  const int64_t* frame_off;  // ragged input: first frame of utterance u in a packed [sum T, V] buffer (or NULL)
  const int32_t* ready;     // streamed input: number of utterances whose logits have landed (or NULL)
  int32_t ready_chunk;      // utterances per host->device chunk
  long long ready_timeout;  // cycles a thread group waits for its chunk before the launch gives up
  int32_t B;
  int32_t* out_n;
  double* out_logit;
  double* out_comb;
  uint8_t* out_tokens;
  int32_t* out_len;
  int32_t* out_status;
  int32_t* out_frames;      // [B, n_best, max_words, 2] or NULL
  int32_t* out_nwords;      // [B, n_best]
  int32_t max_words;
  unsigned long long* stats;
  uint8_t* scratch;
  unsigned long long slot_bytes;
  uint32_t node_cap, bnd_cap, ch_size, outs_cap, wf_cap, hist_cap;
  int32_t* work;
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

__host__ __device__ inline size_t slot_layout(uint32_t node_cap, uint32_t bnd_cap, uint32_t ch_size,
                                              uint32_t outs_cap, uint32_t wf_cap, uint32_t hist_cap, size_t off[8]) {
  size_t o = 0;
  off[0] = o; o = align16(o + (size_t)node_cap * 4);       // node_parent
  off[1] = o; o = align16(o + (size_t)node_cap * 4);       // node_info
  off[2] = o; o = align16(o + (size_t)ch_size * 4);        // row sums of the utterance being classified [T_max]
  off[3] = o; o = align16(o + (size_t)wf_cap * sizeof(FrameRec));  // word-frame records (0 without word frames)
  off[4] = o; o = align16(o + (size_t)bnd_cap * sizeof(BndRec));
  off[5] = o; o = align16(o + (size_t)outs_cap * 16);      // overflow candidates: key, logit
  off[6] = o; o = align16(o + (size_t)outs_cap * 16);      // overflow candidates: order, aux, child, info
  off[7] = o; o = align16(o + (size_t)hist_cap * sizeof(HistRec));  // prune_history records (0 when off)
  return o;
}

// One thread group (= one CTA of NT threads) decodes one utterance at a time and then
// fetches the next from a global counter; `order` lets the host hand out long
// utterances first so the tail of the batch is short.
// minimum CTAs per SM the register allocation must allow: what shared memory permits
template <int NT, int BW, int OUTC, bool FRAMES>
constexpr int min_ctas() {
  constexpr int by_smem = (int)(233472 / (sizeof(GroupShared<BW, OUTC, FRAMES>) + 1024));  // 228 KB per SM, 1 KB reserved per CTA
  constexpr int by_threads = 2048 / NT;
  constexpr int by_regs = 65536 / (NT * 64);  // never ask for fewer than 64 registers per thread
  constexpr int m = by_smem < by_threads ? by_smem : by_threads;
  return m < 1 ? 1 : (m < by_regs ? m : by_regs);
}

template <int NT, int BW, int OUTC, bool FRAMES, bool STATS>
__global__ void __launch_bounds__(NT, min_ctas<NT, BW, OUTC, FRAMES>()) beam_search_kernel(const __grid_constant__ BeamLaunch L) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using Dec = BeamDecoder<NT, BW, OUTC, FRAMES, STATS>;
  typename Dec::Sm& sm = *reinterpret_cast<typename Dec::Sm*>(smem_raw);
  const uint32_t slot = blockIdx.x;
  SlotScratch sc;
  {
    size_t off[8];
    slot_layout(L.node_cap, L.bnd_cap, L.ch_size, L.outs_cap, L.wf_cap, L.hist_cap, off);
    uint8_t* base = L.scratch + (size_t)slot * L.slot_bytes;
    sc.node_parent = reinterpret_cast<uint32_t*>(base + off[0]);
    sc.node_info = reinterpret_cast<uint32_t*>(base + off[1]);
    sc.rowsum = reinterpret_cast<float*>(base + off[2]);
    sc.wf = reinterpret_cast<FrameRec*>(base + off[3]);
    sc.wf_cap = L.wf_cap;
    sc.bnd = reinterpret_cast<BndRec*>(base + off[4]);
    sc.hist = reinterpret_cast<HistRec*>(base + off[7]);
    sc.outs_g.key = reinterpret_cast<unsigned long long*>(base + off[5]);
    sc.outs_g.logit = reinterpret_cast<double*>(base + off[5] + (size_t)L.outs_cap * 8);
    sc.outs_g.order = reinterpret_cast<uint32_t*>(base + off[6]);
    sc.outs_g.aux = sc.outs_g.order + L.outs_cap;
    sc.outs_g.child = sc.outs_g.aux + L.outs_cap;
    sc.outs_g.info = sc.outs_g.child + L.outs_cap;
    sc.node_cap = L.node_cap;
    sc.bnd_cap = L.bnd_cap;
    sc.outs_cap = L.outs_cap;
  }
  for (;;) {
    if (threadIdx.x == 0) sm.utt = atomicAdd(L.work, 1);
    group_sync<NT>();
    const int i = sm.utt;
    if (i >= L.B) break;
    const int u = L.order ? L.order[i] : i;
    if (L.ready != nullptr) {
      // streamed input: wait until the copy stream has delivered this utterance's chunk
      // (bounded: a copier that never delivers must not hang the device -- about 20 s)
      if (threadIdx.x == 0) {
        const int need = min(L.B, (u / L.ready_chunk + 1) * L.ready_chunk);
        const long long t0 = clock64();
        int ok = 1;
        volatile int32_t* gave_up = L.work + 1;  // set by the first group that timed out: nobody waits again
        while (*reinterpret_cast<const volatile int32_t*>(L.ready) < need) {
          __nanosleep(500);
          if (*gave_up || clock64() - t0 > L.ready_timeout) { ok = 0; *gave_up = 1; break; }
        }
        __threadfence();
        sm.status = ok;
      }
      group_sync<NT>();
      const int arrived = sm.status;
      group_sync<NT>();
      if (!arrived) {
        if (threadIdx.x == 0) { L.out_n[u] = 0; L.out_status[u] = CORAL_ECUDA; }
        continue;
      }
    }
    UttIO io;
    io.logits = L.logits + (L.frame_off ? (size_t)L.frame_off[u] : (size_t)u * L.P.T_max) * L.P.V;
    io.T = L.lengths[u];
    io.out_n = L.out_n + u;
    io.out_logit = L.out_logit + (size_t)u * L.P.n_best;
    io.out_comb = L.out_comb + (size_t)u * L.P.n_best;
    io.out_tokens = L.out_tokens + (size_t)u * L.P.n_best * L.P.T_max;
    io.out_len = L.out_len + (size_t)u * L.P.n_best;
    io.out_status = L.out_status + u;
    io.out_frames = FRAMES ? L.out_frames + (size_t)u * L.P.n_best * L.max_words * 2 : nullptr;
    io.out_nwords = FRAMES ? L.out_nwords + (size_t)u * L.P.n_best : nullptr;
    io.max_words = L.max_words;
    io.stats = L.stats;
    Dec::decode(sm, L.lm, L.P, sc, io);
    group_sync<NT>();
  }
}

template <int NT, int BW, int OUTC, bool FRAMES, bool STATS>
static int32_t launch_beam_t(coral_decoder* dec, BeamLaunch& L, int32_t B, cudaStream_t st) {
  using Dec = BeamDecoder<NT, BW, OUTC, FRAMES, STATS>;
  const size_t smem = sizeof(typename Dec::Sm);
  auto kern = beam_search_kernel<NT, BW, OUTC, FRAMES, STATS>;
  // occupancy of this instantiation, queried once per device (the runtime calls are not free
  // and this function sits on the latency path of small batches)
  static int per_sm_cache[64] = {0};
  const int dev_slot = dec->device & 63;
  int per_sm = per_sm_cache[dev_slot];
  if (per_sm == 0) {
    CORAL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CORAL_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
    if (per_sm < 1) return fail(CORAL_ECUDA, "beam kernel does not fit on an SM");
    per_sm_cache[dev_slot] = per_sm;
  }
  const uint32_t want = (uint32_t)std::min<int64_t>((int64_t)B, (int64_t)per_sm * sm_count(dec->device));

  // scratch: worst-case arenas per slot (every frame can add beam_width back-pointer records
  // and beam_width LM boundary records), bounded by a memory budget.
  const uint64_t T = (uint64_t)std::max(1, L.P.T_max);
  const uint64_t bw = (uint64_t)L.P.beam_width;
  uint32_t node_cap = (uint32_t)std::min<uint64_t>(bw * T + 64, 0x7FFFFFFFu);
  const bool need_bnd = L.lm.present || L.P.prune_history;
  uint32_t bnd_cap = need_bnd ? (uint32_t)std::min<uint64_t>(bw * T + 64, (1u << 24) - 1) : 16;
  uint32_t hist_cap = L.P.prune_history ? bnd_cap : 0u;
  uint32_t ch_size = (uint32_t)T + 16;  // floats of row-sum scratch (input classification)
  uint32_t outs_cap = (uint32_t)((bw * (uint64_t)(L.P.V + 1) + 64 + 3) & ~(uint64_t)3);
  uint32_t wf_cap = FRAMES ? (uint32_t)std::min<uint64_t>(bw * T + 64, 0x7FFFFFFFu) : 0u;
  size_t off[8];
  const size_t slot_bytes = slot_layout(node_cap, bnd_cap, ch_size, outs_cap, wf_cap, hist_cap, off);
  std::lock_guard<std::mutex> lock(dec->mu);
  coral_decoder::Scratch& S = dec->scratch[(void*)st];
  // the arena is reused as long as its per-slot capacities cover this launch and it has a slot
  // for every CTA the launch wants (or was already capped by the memory budget); only a
  // (re)allocation queries the free memory
  uint32_t n_slots = want;
  const bool fits = S.d_scratch && S.node_cap >= node_cap && S.bnd_cap >= bnd_cap &&
                    S.ch_size >= ch_size && S.outs_cap >= outs_cap && S.wf_cap >= wf_cap && S.hist_cap >= hist_cap &&
                    (S.n_slots >= want || S.budget_capped);
  if (fits) {
    // reuse the arena with the (larger) capacities it was laid out for
    node_cap = S.node_cap;
    bnd_cap = S.bnd_cap;
    ch_size = S.ch_size;
    outs_cap = S.outs_cap;
    wf_cap = S.wf_cap;
    hist_cap = S.hist_cap;
  } else {
    node_cap = std::max(node_cap, S.node_cap);
    bnd_cap = std::max(bnd_cap, S.bnd_cap);
    ch_size = std::max(ch_size, S.ch_size);
    outs_cap = std::max(outs_cap, S.outs_cap);
    wf_cap = std::max(wf_cap, S.wf_cap);
    hist_cap = std::max(hist_cap, S.hist_cap);
    if (hist_cap) hist_cap = std::max(hist_cap, bnd_cap);  // one history record per boundary record
    const size_t sb = slot_layout(node_cap, bnd_cap, ch_size, outs_cap, wf_cap, hist_cap, off);
    size_t free_b = 0, total_b = 0;
    CORAL_CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
    const size_t budget = std::min<size_t>((size_t)48 << 30, (free_b + S.scratch_bytes) / 2);
    const size_t wanted = std::max(want, S.n_slots);
    n_slots = (uint32_t)std::max<size_t>(1, std::min<size_t>(wanted, budget / sb));
    S.budget_capped = n_slots < wanted;
    // wait for this stream's earlier launches, which may still use the old arena, then rebuild it
    CORAL_CUDA_OK(cudaStreamSynchronize(st));
    if (S.d_scratch) cudaFree(S.d_scratch);
    S.d_scratch = nullptr;
    S.scratch_bytes = 0;
    S.n_slots = 0;
    const size_t bytes = sb * n_slots;
    CORAL_CUDA_OK(cudaMalloc(&S.d_scratch, bytes));
    // nothing to initialise: every arena record is written before it is read
    S.scratch_bytes = bytes;
    S.slot_bytes = sb;
    S.n_slots = n_slots;
    S.node_cap = node_cap;
    S.bnd_cap = bnd_cap;
    S.ch_size = ch_size;
    S.outs_cap = outs_cap;
    S.wf_cap = wf_cap;
    S.hist_cap = hist_cap;
  }
  if (!S.d_work) CORAL_CUDA_OK(cudaMalloc(&S.d_work, 2 * sizeof(int32_t)));  // work counter, give-up flag
  CORAL_CUDA_OK(cudaMemsetAsync(S.d_work, 0, 2 * sizeof(int32_t), st));
  L.scratch = S.d_scratch;
  L.slot_bytes = S.slot_bytes;
  L.node_cap = node_cap;
  L.bnd_cap = bnd_cap;
  L.ch_size = ch_size;
  L.outs_cap = outs_cap;
  L.wf_cap = wf_cap;
  L.hist_cap = hist_cap;
  L.work = S.d_work;
  const uint32_t grid = std::min<uint32_t>(S.n_slots, std::max<uint32_t>(1, want));
  kern<<<grid, NT, smem, st>>>(L);
  CORAL_CUDA_OK(cudaGetLastError());
  return CORAL_OK;
}

// the instantiation with work counters and cycle timers only when the caller passes a stats buffer
template <int NT, int BW, int OUTC, bool FRAMES = false>
static int32_t launch_beam(coral_decoder* dec, BeamLaunch& L, int32_t B, cudaStream_t st) {
  if (L.stats) return launch_beam_t<NT, BW, OUTC, FRAMES, true>(dec, L, B, st);
  return launch_beam_t<NT, BW, OUTC, FRAMES, false>(dec, L, B, st);
}

}  // namespace coral

using namespace coral;

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
    const int rc = build_lexicon(lm->host, n_unigrams >= 0 ? &uni : nullptr, d->lex, err);
    if (rc != 0) { delete d; return fail(rc, err); }
    DeviceGuard g(device);
    const size_t lb = d->lex.lex.size() * sizeof(LexSlot);
    cudaError_t e;
    if ((e = cudaMalloc(&d->d_lex, lb)) != cudaSuccess ||
        (e = cudaMemcpy(d->d_lex, d->lex.lex.data(), lb, cudaMemcpyHostToDevice)) != cudaSuccess) {
      std::string m = std::string("uploading lexicon: ") + cudaGetErrorString(e);
      coral_decoder_free(d);
      return fail(CORAL_ECUDA, m);
    }
    d->device_bytes = lb;
  }
  *out = d;
  return CORAL_OK;
}

int32_t coral_decoder_free(coral_decoder* d) {
  if (!d) return CORAL_OK;
  DeviceGuard g(d->device);
  cudaDeviceSynchronize();
  if (d->d_lex) cudaFree(d->d_lex);
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
    L.lm = make_view(dec->lm->host, dec->lex, dec->lm->d_uni, dec->lm->d_ng, dec->d_lex);
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

  if (out_word_frames_dev) {  // the instantiation that also tracks pyctcdecode's word frames
    if (beam_width <= 32) return launch_beam<32, 32, 128, true>(dec, L, B, st);
    if (beam_width <= 64) return launch_beam<64, 64, 192, true>(dec, L, B, st);
    if (beam_width <= 128) return launch_beam<128, 128, 320, true>(dec, L, B, st);
    if (beam_width <= 256) return launch_beam<256, 256, 640, true>(dec, L, B, st);
    return launch_beam<256, 512, 1280, true>(dec, L, B, st);
  }

  // threads per utterance: CORAL_BEAM_NT overrides the default (tuning knob)
  int nt = 0;
  if (const char* e = getenv("CORAL_BEAM_NT")) nt = atoi(e);
  if (beam_width <= 32) {
    if (nt == 64) return launch_beam<64, 32, 128>(dec, L, B, st);
    return launch_beam<32, 32, 128>(dec, L, B, st);
  }
  if (beam_width <= 64) {
    if (nt == 32) return launch_beam<32, 64, 192>(dec, L, B, st);
    if (nt == 128) return launch_beam<128, 64, 192>(dec, L, B, st);
    return launch_beam<64, 64, 192>(dec, L, B, st);
  }
  // up to 104 beams (pyctcdecode's default is 100): arrays sized so that EIGHT thread groups fit
  // on an SM (27.9 KB of shared memory, 64 registers) against the 128-beam instantiation's six.
  // Resident groups per SM 4 / 5 / 6 / 7 / 8: 22.4 / 19.0 / 17.1 / 15.5 / 15.1 ms per 8192 utterances
  if (beam_width <= 104 && nt == 0) return launch_beam<128, 104, 208>(dec, L, B, st);
  if (beam_width <= 128) {
    if (nt == 32) return launch_beam<32, 128, 320>(dec, L, B, st);
    if (nt == 64) return launch_beam<64, 128, 320>(dec, L, B, st);
    return launch_beam<128, 128, 320>(dec, L, B, st);
  }
  if (beam_width <= 256) {
    if (nt == 128) return launch_beam<128, 256, 640>(dec, L, B, st);
    return launch_beam<256, 256, 640>(dec, L, B, st);
  }
  return launch_beam<256, 512, 1280>(dec, L, B, st);
}

}  // extern "C"
