// Persistent beam-search kernel and its launch (shared by beam.cu -- text-only instantiations --
// and beam_frames.cu -- the instantiations that also track pyctcdecode's word frames; two
// translation units so that nvcc compiles them side by side).
#pragma once
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
  int32_t* work;        // [4]: work counter, give-up flag, utterances flagged for the heavy kernel, its work counter
  const uint8_t* heavy_flag;  // [B] 1 = the utterance belongs to the heavy kernel (NULL: no split)
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

__host__ __device__ inline size_t slot_layout(uint32_t node_cap, uint32_t bnd_cap, uint32_t ch_size,
                                              uint32_t outs_cap, uint32_t wf_cap, uint32_t hist_cap, size_t off[7]) {
  size_t o = 0;
  off[0] = o; o = align16(o + (size_t)node_cap * 4);       // node_parent
  off[1] = o; o = align16(o + (size_t)node_cap * 4);       // node_info
  off[2] = o; o = align16(o + (size_t)ch_size * 4);        // row sums of the utterance being classified [T_max]
  off[3] = o; o = align16(o + (size_t)wf_cap * sizeof(FrameRec));  // word-frame records (0 without word frames)
  off[4] = o; o = align16(o + (size_t)bnd_cap * sizeof(BndRec));
  off[5] = o; o = align16(o + (size_t)outs_cap * 16);      // overflow candidates: key, logit
  // overflow candidates: order, aux, child, info; then (addressed from info, see SlotScratch) the
  // heavy-frame arrays hv_sorted u32 [outs_cap], hv_masks u64 [1024], hv_bin u8 [outs_cap] and the prune_history records
  off[6] = o;
  o = o + (size_t)outs_cap * 16 + (size_t)outs_cap * 4 + SlotScratch::kHvMaskBytes + (((size_t)outs_cap + 15) & ~(size_t)15);
  o = align16(o + (size_t)hist_cap * sizeof(HistRec));
  return o;
}

// One thread group (= one CTA of NT threads) decodes one utterance at a time and then
// fetches the next from a global counter; `order` lets the host hand out long
// utterances first so the tail of the batch is short.
// minimum CTAs per SM the register allocation must allow: what shared memory permits
template <int NT, int BW, int OUTC, bool FRAMES>
constexpr int min_ctas() {
  constexpr int by_smem = (int)(233472 / (sizeof(GroupShared<BW, OUTC, FRAMES, false, false>) + 1024));  // 228 KB per SM, 1 KB reserved per CTA
  constexpr int by_threads = 2048 / NT;
  constexpr int by_regs = 65536 / (NT * 64);  // never ask for fewer than 64 registers per thread
  constexpr int m = by_smem < by_threads ? by_smem : by_threads;
  return m < 1 ? 1 : (m < by_regs ? m : by_regs);
}

// Which utterances keep many tokens per frame? One warp per utterance samples four frames (first,
// last and two in between), takes the log-softmax the plain way and counts the tokens that pass
// token_min_logp; above kHeavyKept per frame on average the utterance is flagged and queued for
// the heavy kernel. A heuristic that only picks WHICH kernel decodes the utterance: both kernels
// return the same beams, so its floating-point details are irrelevant. Runs in front of the lean
// kernel (about 10 us for 8192 utterances); not used with streamed input (the frames are not
// there yet), where everything stays with the lean kernel.
constexpr int kHeavyKept = 24;
static __global__ void __launch_bounds__(256)
classify_heavy_kernel(const float* __restrict__ logits, const int64_t* __restrict__ frame_off,
                      const int32_t* __restrict__ lengths, int B, int T_max, int V, float token_min_logp,
                      uint8_t* __restrict__ flags, int32_t* __restrict__ n_heavy) {
  const int lane = threadIdx.x & 31;
  const int u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (u >= B) return;
  const int T = lengths[u];
  const float* base = logits + (frame_off ? (size_t)frame_off[u] : (size_t)u * T_max) * V;
  // all loads first: with the logits in pinned host memory each one is a PCIe round trip
  const int nf = T < 4 ? (T < 0 ? 0 : T) : 4;
  float x0[4], x1[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int t = T <= 4 ? k : (int)(((long long)k * (T - 1)) / 3);
    const float* row = base + (size_t)t * V;
    x0[k] = (k < nf && lane < V) ? row[lane] : -INFINITY;
    x1[k] = (k < nf && lane + 32 < V) ? row[lane + 32] : -INFINITY;
  }
  int kept = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float mx = fmaxf(x0[k], x1[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (!isfinite(mx)) mx = 0.0f;
    float se = (lane < V ? expf(x0[k] - mx) : 0.0f) + (lane + 32 < V ? expf(x1[k] - mx) : 0.0f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    const float lse = logf(se) + mx;
    int c = (lane < V && x0[k] - lse >= token_min_logp) + (lane + 32 < V && x1[k] - lse >= token_min_logp);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    kept += k < nf ? c : 0;
  }
  if (lane == 0) {
    const bool heavy = nf > 0 && kept > nf * kHeavyKept;
    flags[u] = heavy ? 1 : 0;
    if (heavy) atomicAdd(n_heavy, 1);
  }
}

#ifndef CORAL_HEAVY_CTA_DIV
#define CORAL_HEAVY_CTA_DIV 2
#endif
// MODE 0: the only kernel of the launch (instrumented builds): everything, heavy frames included.
// MODE 1: the lean kernel: no heavy-frame code; an utterance whose first frames keep many tokens is
//         left untouched and queued for
// MODE 2: the heavy kernel, which decodes the queued utterances (it runs right behind the lean one
//         on the same stream and shares its scratch arenas).
// the heavy kernel trades resident thread groups for registers: its per-item bound computation
// keeps a dozen per-prefix values live, and it is throughput- not latency-bound
template <int NT, int BW, int OUTC, bool FRAMES, int MODE>
constexpr int kernel_min_ctas() {
  constexpr int m = min_ctas<NT, BW, OUTC, FRAMES>();
  // (only where there are thread groups to spare: the wide-beam instantiations hold 1-3 per SM)
  return (MODE == 2 && m >= 6) ? m / CORAL_HEAVY_CTA_DIV : m;
}

template <int NT, int BW, int OUTC, bool FRAMES, bool STATS, int MODE>
__global__ void __launch_bounds__(NT, kernel_min_ctas<NT, BW, OUTC, FRAMES, MODE>()) beam_search_kernel(const __grid_constant__ BeamLaunch L) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using Dec = BeamDecoder<NT, BW, OUTC, FRAMES, STATS, MODE != 1>;
  typename Dec::Sm& sm = *reinterpret_cast<typename Dec::Sm*>(smem_raw);
  const uint32_t slot = blockIdx.x;
  SlotScratch sc;
  {
    size_t off[7];
    slot_layout(L.node_cap, L.bnd_cap, L.ch_size, L.outs_cap, L.wf_cap, L.hist_cap, off);
    uint8_t* base = L.scratch + (size_t)slot * L.slot_bytes;
    sc.node_parent = reinterpret_cast<uint32_t*>(base + off[0]);
    sc.node_info = reinterpret_cast<uint32_t*>(base + off[1]);
    sc.rowsum = reinterpret_cast<float*>(base + off[2]);
    sc.wf = reinterpret_cast<FrameRec*>(base + off[3]);
    sc.wf_cap = L.wf_cap;
    sc.bnd = reinterpret_cast<BndRec*>(base + off[4]);
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
    if (MODE == 2 && L.work[2] == 0) break;  // nothing was flagged: the heavy kernel has no work
    if (threadIdx.x == 0) sm.utt = atomicAdd(L.work + (MODE == 2 ? 3 : 0), 1);
    group_sync<NT>();
    const int i = sm.utt;
    group_sync<NT>();  // the skips below loop straight back to the next claim: everyone has read this one
    if (i >= L.B) break;
    const int u = L.order ? L.order[i] : i;
    // many kept tokens per frame (flat posteriors, a loose token_min_logp): that utterance belongs to
    // the heavy kernel, which runs right behind the lean one and walks the batch in the same
    // (longest first) order (classify_heavy_kernel decided)
    if (MODE == 1 && L.heavy_flag != nullptr && L.heavy_flag[u]) continue;
    if (MODE == 2 && !L.heavy_flag[u]) continue;
    if (MODE != 2 && L.ready != nullptr) {
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

template <int NT, int BW, int OUTC, bool FRAMES, bool STATS, int MODE>
static int32_t kernel_occupancy(coral_decoder* dec, int* per_sm_out) {
  using Dec = BeamDecoder<NT, BW, OUTC, FRAMES, STATS, MODE != 1>;
  const size_t smem = sizeof(typename Dec::Sm);
  auto kern = beam_search_kernel<NT, BW, OUTC, FRAMES, STATS, MODE>;
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
  *per_sm_out = per_sm;
  return CORAL_OK;
}

template <int NT, int BW, int OUTC, bool FRAMES, bool STATS>
static int32_t launch_beam_t(coral_decoder* dec, BeamLaunch& L, int32_t B, cudaStream_t st) {
  // instrumented launches run ONE kernel that contains everything; production launches run the
  // lean kernel and, right behind it, the heavy kernel over whatever the lean one left to it
  constexpr int kFirst = STATS ? 0 : 1;
  using Dec = BeamDecoder<NT, BW, OUTC, FRAMES, STATS, kFirst != 1>;
  const size_t smem = sizeof(typename Dec::Sm);
  int per_sm = 0, per_sm_heavy = 0;
  {
    const int32_t rc = kernel_occupancy<NT, BW, OUTC, FRAMES, STATS, kFirst>(dec, &per_sm);
    if (rc != CORAL_OK) return rc;
    if (!STATS) {
      const int32_t rc2 = kernel_occupancy<NT, BW, OUTC, FRAMES, STATS, 2>(dec, &per_sm_heavy);
      if (rc2 != CORAL_OK) return rc2;
    }
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
  size_t off[7];
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
  // work counters (4 ints: lean counter, give-up flag, flagged utterances, heavy counter) + one flag byte per utterance
  const size_t work_ints = 4 + ((size_t)B + 3) / 4;
  if (S.work_cap < work_ints) {
    CORAL_CUDA_OK(cudaStreamSynchronize(st));
    if (S.d_work) cudaFree(S.d_work);
    S.d_work = nullptr;
    S.work_cap = 0;
    const size_t cap = std::max<size_t>(work_ints, 8192);
    CORAL_CUDA_OK(cudaMalloc(&S.d_work, cap * sizeof(int32_t)));
    S.work_cap = cap;
  }
  CORAL_CUDA_OK(cudaMemsetAsync(S.d_work, 0, 4 * sizeof(int32_t), st));
  L.scratch = S.d_scratch;
  L.slot_bytes = S.slot_bytes;
  L.node_cap = node_cap;
  L.bnd_cap = bnd_cap;
  L.ch_size = ch_size;
  L.outs_cap = outs_cap;
  L.wf_cap = wf_cap;
  L.hist_cap = hist_cap;
  L.work = S.d_work;
  L.heavy_flag = nullptr;
  // heavy utterances are picked out in front of the lean kernel -- unless the logits are still
  // arriving (streamed input), or the vocabulary is wider than the classifier's two tokens per lane
  const bool split = !STATS && L.ready == nullptr && L.P.V <= 64;
  if (split) {
    uint8_t* flags = reinterpret_cast<uint8_t*>(S.d_work + 4);
    classify_heavy_kernel<<<(unsigned)((B + 7) / 8), 256, 0, st>>>(L.logits, L.frame_off, L.lengths, B, L.P.T_max, L.P.V,
                                                                  L.P.token_min_logp, flags, L.work + 2);
    CORAL_CUDA_OK(cudaGetLastError());
    L.heavy_flag = flags;
  }
  const uint32_t grid = std::min<uint32_t>(S.n_slots, std::max<uint32_t>(1, want));
  beam_search_kernel<NT, BW, OUTC, FRAMES, STATS, kFirst><<<grid, NT, smem, st>>>(L);
  CORAL_CUDA_OK(cudaGetLastError());
  if (split) {
    using DecH = BeamDecoder<NT, BW, OUTC, FRAMES, STATS, true>;
    const uint32_t want_h = (uint32_t)std::min<int64_t>((int64_t)B, (int64_t)per_sm_heavy * sm_count(dec->device));
    const uint32_t grid_h = std::min<uint32_t>(S.n_slots, std::max<uint32_t>(1, want_h));
    beam_search_kernel<NT, BW, OUTC, FRAMES, STATS, 2><<<grid_h, NT, sizeof(typename DecH::Sm), st>>>(L);
    CORAL_CUDA_OK(cudaGetLastError());
  }
  return CORAL_OK;
}

// the instantiation with work counters and cycle timers only when the caller passes a stats buffer
template <int NT, int BW, int OUTC, bool FRAMES = false>
static int32_t launch_beam(coral_decoder* dec, BeamLaunch& L, int32_t B, cudaStream_t st) {
  if (L.stats) return launch_beam_t<NT, BW, OUTC, FRAMES, true>(dec, L, B, st);
  return launch_beam_t<NT, BW, OUTC, FRAMES, false>(dec, L, B, st);
}

// defined in beam_frames.cu: the word-frame instantiations by beam width
int32_t launch_beam_frames(coral_decoder* dec, BeamLaunch& L, int32_t B, int32_t beam_width, cudaStream_t st);

}  // namespace coral
