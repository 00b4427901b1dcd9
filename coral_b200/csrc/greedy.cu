// C-ABI: greedy CTC decode -- per-frame argmax (HBM-bound streaming kernel) and the
// repeat-collapse / blank-drop compaction.
//
// Replaces np.argmax(predictions, axis=-1) at R:src/coral/compute_metrics.py:62-68 (with
// its "-100 row -> pad" fix-up) and the itertools.groupby collapse + pad filtering of
// Wav2Vec2CTCTokenizer.convert_tokens_to_string
// (HF:models/wav2vec2/tokenization_wav2vec2.py:296-357, :410-459). Id -> string mapping
// stays on the host (coral_b200/greedy.py).
//
// Roofline: the argmax kernel reads every logit once (T*V*4 bytes per
// utterance) and writes T*4 bytes; nothing is re-read. Tiles of 128 frames are brought into
// shared memory by ONE bulk asynchronous copy (TMA, cp.async.bulk + mbarrier) issued by one
// thread, so the other threads spend their instructions on the row scans only; about nine
// CTAs per SM keep enough tiles in flight to cover the HBM latency.
#include <algorithm>
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace coral {

constexpr int kTileFrames = 128;  // frames per tile (= threads per CTA)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem)), "l"(gmem));
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem)), "l"(gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}

// Fallback copy of one tile (n_floats contiguous floats) for the tiles a 16-byte-granular bulk
// copy must not touch (it would read a few bytes outside the caller's buffer).
__device__ __forceinline__ void issue_tile(const float* __restrict__ src, float* stage, int n_floats) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(src);
  if ((a & 15) == 0) {
    const int n16 = n_floats >> 2;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) cp_async16(stage + 4 * i, src + 4 * i);
    for (int e = (n16 << 2) + threadIdx.x; e < n_floats; e += blockDim.x) cp_async4(stage + e, src + e);
  } else if ((a & 7) == 0) {
    const int n8 = n_floats >> 1;
    for (int i = threadIdx.x; i < n8; i += blockDim.x) cp_async8(stage + 2 * i, src + 2 * i);
    for (int e = (n8 << 1) + threadIdx.x; e < n_floats; e += blockDim.x) cp_async4(stage + e, src + e);
  } else {
    for (int e = threadIdx.x; e < n_floats; e += blockDim.x) cp_async4(stage + e, src + e);
  }
  asm volatile("cp.async.commit_group;");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// numpy.argmax of one row in shared memory: first maximum wins, the first NaN beats everything.
__device__ __forceinline__ int row_argmax(const float* row, int V, float& best_out) {
  float best;
  int id;
  bool nan;
  if ((V & 1) == 0 && (reinterpret_cast<uintptr_t>(row) & 7) == 0) {
    // two independent chains over the even and the odd elements, 8-byte shared loads
    const float2* r2 = reinterpret_cast<const float2*>(row);
    float2 x = r2[0];
    float ba = x.x, bb = x.y;
    int ia = 0, ib = 1;
    nan = (x.x != x.x) | (x.y != x.y);
#pragma unroll 4
    for (int v = 1; v < (V >> 1); ++v) {
      x = r2[v];
      nan |= (x.x != x.x) | (x.y != x.y);
      if (x.x > ba) { ba = x.x; ia = 2 * v; }
      if (x.y > bb) { bb = x.y; ib = 2 * v + 1; }
    }
    const bool take_b = bb > ba || (bb == ba && ib < ia);
    best = take_b ? bb : ba;
    id = take_b ? ib : ia;
  } else {
    best = row[0];
    id = 0;
    nan = best != best;
#pragma unroll 4
    for (int v = 1; v < V; ++v) {
      const float x = row[v];
      nan |= x != x;
      if (x > best) { best = x; id = v; }
    }
  }
  if (nan) {  // rare: numpy returns the first NaN
    for (int v = 0; v < V; ++v)
      if (row[v] != row[v]) { best = row[v]; id = v; break; }
  }
  best_out = best;
  return id;
}

// Persistent grid over (utterance, 128-frame block) tiles; blocks past an utterance's length
// are skipped. Thread f scans row f of the tile from shared memory.
__global__ void __launch_bounds__(kTileFrames)
ctc_argmax_kernel(const float* __restrict__ logits, const int32_t* __restrict__ lengths, int B, int T_max, int V,
                  int blank_id, int pad_fixup, int32_t* __restrict__ out_ids) {
  extern __shared__ __align__(128) unsigned char stage[];
  __shared__ __align__(8) unsigned long long bar;
  const int tpu = (T_max + kTileFrames - 1) / kTileFrames;
  const long long n_tiles = (long long)B * tpu;
  const uintptr_t buf_lo = reinterpret_cast<uintptr_t>(logits);
  const uintptr_t buf_hi = buf_lo + (size_t)B * T_max * V * sizeof(float);
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  unsigned phase = 0;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int u = (int)(tile / tpu);
    const int t0 = (int)(tile % tpu) * kTileFrames;
    const int T = lengths ? lengths[u] : T_max;
    const int nfr = T - t0 < kTileFrames ? T - t0 : kTileFrames;
    if (nfr <= 0) continue;  // uniform
    const float* src = logits + ((size_t)u * T_max + t0) * V;
    const uintptr_t a = reinterpret_cast<uintptr_t>(src);
    const uintptr_t a0 = a & ~(uintptr_t)15;
    unsigned lead = (unsigned)(a - a0);
    const unsigned bytes = (lead + (unsigned)nfr * V * 4u + 15u) & ~15u;
    if (a0 >= buf_lo && a0 + bytes <= buf_hi) {
      if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, bytes);
        bulk_g2s(stage, reinterpret_cast<const void*>(a0), bytes, &bar);
      }
      mbar_wait(&bar, phase);
      phase ^= 1u;
    } else {
      lead = 0;
      issue_tile(src, reinterpret_cast<float*>(stage), nfr * V);
      __syncthreads();
    }
    const int f = threadIdx.x;
    if (f < nfr) {
      const float* row = reinterpret_cast<const float*>(stage + lead) + (size_t)f * V;
      float best;
      int id = row_argmax(row, V, best);
      if (pad_fixup && best == -100.0f) {  // "all -100 row -> pad" (R:src/coral/compute_metrics.py:66)
        bool all_m100 = true;
        for (int v = 0; v < V; ++v) all_m100 &= row[v] == -100.0f;
        if (all_m100) id = blank_id;
      }
      out_ids[(size_t)u * T_max + t0 + f] = id;
    }
    __syncthreads();  // every row was read before the next tile overwrites the stage
  }
}

// ---------------------------------------------------------------------------------------------
// Fused greedy decode: argmax + repeat-collapse + blank-drop in ONE kernel, no ids round trip.
//
// Every CTA walks a sequence of tiles (utterance, 128-frame block); utterances are handed out by a
// global counter, so the sequence is only known to the CTA's producer thread, which runs
// kStages - 1 tiles ahead: it writes the tile descriptor into the stage's slot, issues one bulk
// asynchronous copy (TMA) for the tile and moves on -- there is always a tile in flight while the
// 128 consumer threads scan the rows of the current one (thread f = frame f). The collapse runs
// tile by tile inside the utterance: keep[f] = id != blank && (first frame || id != previous id),
// an exclusive scan over the CTA (ballots + four warp totals) gives the output slot, the id of the
// tile's last frame and the running output length are carried to the next tile.
constexpr int kStages = 3;

struct TileDesc {
  int u;        // utterance, -1 = no more work
  int t0;       // first frame of the tile
  int nfr;      // frames in the tile
  int last;     // 1 = last tile of the utterance
  unsigned lead;  // bytes between the 16-byte aligned copy start and the first row
  int bulk;     // 1 = delivered by the bulk copy (wait on the mbarrier), 0 = cp.async fallback
};

__global__ void __launch_bounds__(kTileFrames)
ctc_greedy_fused_kernel(const float* __restrict__ logits, const int32_t* __restrict__ lengths, int B, int T_max, int V,
                        int blank_id, int pad_fixup, int32_t* __restrict__ out_ids, int32_t* __restrict__ out_tokens,
                        int32_t* __restrict__ out_lens, int* __restrict__ work, unsigned stage_bytes) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long full[kStages];
  __shared__ TileDesc desc[kStages];
  __shared__ int s_ids[2][kTileFrames];
  __shared__ int s_wtot[2][kTileFrames / 32];
  const uintptr_t buf_lo = reinterpret_cast<uintptr_t>(logits);
  const uintptr_t buf_hi = buf_lo + (size_t)B * T_max * V * sizeof(float);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0)
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
  __syncthreads();

  // producer state (thread 0 only). Utterances are claimed from the global counter two claims ahead
  // of their use: the atomic for claim k + 2 and the length load of claim k + 1 are issued when
  // utterance k starts, so neither round trip (about a microsecond each, as long as a whole
  // utterance takes to stream) is ever waited for. One utterance per claim keeps the tail short:
  // a CTA never sits on more than three unprocessed utterances.
  constexpr int kGrab = 1;
  int pu = -1, pT = 0, pt0 = 0;
  bool drained = false;
  int b_u0 = 0, b_i = kGrab;                // current batch: first utterance, next index in it
  int b_len[kGrab] = {0};
  int n_u0 = 0, n_len[kGrab] = {0};         // next batch (lengths already requested)
  int nn_u0 = 0;                            // the batch after that (just claimed)
  auto load_lens = [&](int u0, int* len) {
#pragma unroll
    for (int j = 0; j < kGrab; ++j) {
      const int u = u0 + j;
      int T = u < B ? (lengths ? lengths[u] : T_max) : 0;
      len[j] = T < 0 ? 0 : (T > T_max ? T_max : T);
    }
  };
  if (tid == 0) {
    b_u0 = atomicAdd(work, kGrab);
    n_u0 = atomicAdd(work, kGrab);
    nn_u0 = atomicAdd(work, kGrab);
    load_lens(b_u0, b_len);
    load_lens(n_u0, n_len);
    b_i = 0;
  }
  auto produce = [&](int s) {
    // next tile of the current utterance, or the first tile of the next utterance with frames
    if (!drained && (pu < 0 || pt0 >= pT)) {
      for (;;) {
        if (b_i == kGrab) {  // rotate the batches and keep two claims in flight
          b_u0 = n_u0;
#pragma unroll
          for (int j = 0; j < kGrab; ++j) b_len[j] = n_len[j];
          n_u0 = nn_u0;
          if (n_u0 < B) { load_lens(n_u0, n_len); nn_u0 = atomicAdd(work, kGrab); }
          b_i = 0;
        }
        pu = b_u0 + b_i;
        if (pu >= B) { drained = true; break; }
        pT = b_len[0];
#pragma unroll
        for (int j = 1; j < kGrab; ++j) pT = b_i == j ? b_len[j] : pT;
        ++b_i;
        pt0 = 0;
        if (pT > 0) break;
        out_lens[pu] = 0;  // an empty utterance has no tile
      }
    }
    TileDesc d;
    if (drained) {
      d.u = -1; d.t0 = 0; d.nfr = 0; d.last = 1; d.lead = 0; d.bulk = 0;
      desc[s] = d;
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[s])) : "memory");
      return;
    }
    d.u = pu; d.t0 = pt0;
    d.nfr = pT - pt0 < kTileFrames ? pT - pt0 : kTileFrames;
    d.last = pt0 + d.nfr >= pT;
    const float* src = logits + ((size_t)pu * T_max + pt0) * V;
    const uintptr_t a = reinterpret_cast<uintptr_t>(src);
    const uintptr_t a0 = a & ~(uintptr_t)15;
    d.lead = (unsigned)(a - a0);
    const unsigned bytes = (d.lead + (unsigned)d.nfr * V * 4u + 15u) & ~15u;
    d.bulk = (a0 >= buf_lo && a0 + bytes <= buf_hi) ? 1 : 0;
    if (!d.bulk) d.lead = 0;
    desc[s] = d;
    if (d.bulk) {
      mbar_expect_tx(&full[s], bytes);
      bulk_g2s(smem + (size_t)s * stage_bytes, reinterpret_cast<const void*>(a0), bytes, &full[s]);
    } else {
      // a tile the 16-byte granular bulk copy must not touch: the consumers copy it themselves
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[s])) : "memory");
    }
    pt0 += d.nfr;
  };
  if (tid == 0)
    for (int s = 0; s < kStages - 1; ++s) produce(s);

  int carry = -1;   // id of the utterance's previous frame (uniform across the CTA)
  int base = 0;     // tokens written so far for the utterance
  unsigned phase_bits = 0;
  for (int k = 0;; ++k) {
    const int s = k % kStages;
    if (tid == 0) produce((k + kStages - 1) % kStages);  // that stage was released by the barrier of tile k - 1
    mbar_wait(&full[s], (phase_bits >> s) & 1u);
    phase_bits ^= 1u << s;
    const TileDesc d = desc[s];
    if (d.u < 0) break;
    unsigned char* stage = smem + (size_t)s * stage_bytes;
    if (!d.bulk) {
      issue_tile(logits + ((size_t)d.u * T_max + d.t0) * V, reinterpret_cast<float*>(stage), d.nfr * V);
      __syncthreads();
    }
    const int par = k & 1;
    int id = blank_id;
    if (tid < d.nfr) {
      const float* row = reinterpret_cast<const float*>(stage + d.lead) + (size_t)tid * V;
      float best;
      id = row_argmax(row, V, best);
      if (pad_fixup && best == -100.0f) {  // "all -100 row -> pad" (R:src/coral/compute_metrics.py:66)
        bool all_m100 = true;
        for (int v = 0; v < V; ++v) all_m100 &= row[v] == -100.0f;
        if (all_m100) id = blank_id;
      }
      if (out_ids) out_ids[(size_t)d.u * T_max + d.t0 + tid] = id;
    }
    s_ids[par][tid] = id;
    // the previous frame's id: the neighbour lane, the previous warp's last lane, or the carry
    int prev = __shfl_up_sync(0xffffffffu, id, 1);
    bool keep = false;
    unsigned m = 0;
    // warps need each other's last id and totals: one barrier per tile, which also releases the
    // stage (every row has been read) for the producer
    __syncthreads();
    if (lane == 0) prev = warp == 0 ? carry : s_ids[par][tid - 1];
    keep = tid < d.nfr && id != blank_id && ((d.t0 == 0 && tid == 0) || id != prev);
    m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_wtot[par][warp] = __popc(m);
    // second, cheap barrier for the warp totals (4 ints)
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kTileFrames / 32; ++w) {
      const int c = s_wtot[par][w];
      before += w < warp ? c : 0;
      total += c;
    }
    if (keep) out_tokens[(size_t)d.u * T_max + base + before + __popc(m & ((1u << lane) - 1u))] = id;
    carry = s_ids[par][d.nfr - 1];
    base += total;
    if (d.last) {
      if (tid == 0) out_lens[d.u] = base;
      base = 0;
      carry = -1;
    }
  }
  // the last CTA to leave rewinds the claim counter for the next launch on this stream (saves a
  // memset launch in front of a ~100 us kernel)
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(work + 1, 1) == (int)gridDim.x - 1) { work[0] = 0; work[1] = 0; __threadfence(); }
  }
}

// One warp per utterance: keep[t] = id != blank && (!group || t == 0 || id != ids[t-1]),
// compacted with ballots (no block barriers). Safe in place: a chunk is loaded completely
// before anything is stored, and stores never pass the read cursor.
constexpr int kCollapseWarps = 8;
__global__ void __launch_bounds__(kCollapseWarps * 32)
ctc_collapse_kernel(const int32_t* ids, const int32_t* __restrict__ lengths, int B, int T_max, int blank_id,
                    int group_tokens, int32_t* out_tokens, int32_t* __restrict__ out_lens) {
  const int lane = threadIdx.x & 31;
  const int u = blockIdx.x * kCollapseWarps + (threadIdx.x >> 5);
  if (u >= B) return;
  const int T = lengths ? lengths[u] : T_max;
  const int32_t* src = ids + (size_t)u * T_max;
  int32_t* dst = out_tokens + (size_t)u * T_max;
  int base = 0;
  int carry = -1;  // id of the frame before the current chunk
  for (int c0 = 0; c0 < T; c0 += 128) {
    int id[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int t = c0 + 32 * k + lane;
      id[k] = t < T ? src[t] : blank_id;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int t = c0 + 32 * k + lane;
      int prev = __shfl_up_sync(0xffffffffu, id[k], 1);
      if (lane == 0) prev = carry;
      const bool keep = t < T && id[k] != blank_id && (!group_tokens || t == 0 || id[k] != prev);
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) dst[base + __popc(m & ((1u << lane) - 1u))] = id[k];
      base += __popc(m);
      carry = __shfl_sync(0xffffffffu, id[k], 31);
    }
  }
  if (lane == 0) out_lens[u] = base;
}

}  // namespace coral

using namespace coral;

extern "C" {

int32_t coral_ctc_collapse(const int32_t* ids_dev, const int32_t* lengths_dev, int32_t B, int32_t T_max,
                           int32_t blank_id, int32_t group_tokens, int32_t* out_tokens_dev, int32_t* out_lens_dev,
                           void* stream) {
  if (B < 0 || T_max < 0) return fail(CORAL_EARG, "negative batch or frame count");
  if (B == 0) return CORAL_OK;
  if (!ids_dev || !out_tokens_dev || !out_lens_dev) return fail(CORAL_EARG, "coral_ctc_collapse: null buffer");
  ctc_collapse_kernel<<<(unsigned)((B + kCollapseWarps - 1) / kCollapseWarps), kCollapseWarps * 32, 0,
                        (cudaStream_t)stream>>>(ids_dev, lengths_dev, B, T_max, blank_id, group_tokens, out_tokens_dev,
                                                out_lens_dev);
  CORAL_CUDA_OK(cudaGetLastError());
  return CORAL_OK;
}

// per-(device, stream) work counters of the fused kernel
static std::mutex g_greedy_mu;
static std::map<std::pair<int, void*>, int*> g_greedy_work;

int32_t coral_ctc_greedy(const float* logits_dev, const int32_t* lengths_dev, int32_t B, int32_t T_max, int32_t V,
                         int32_t blank_id, int32_t pad_fixup, int32_t* out_ids_dev, int32_t* out_tokens_dev,
                         int32_t* out_lens_dev, void* stream) {
  if (B < 0 || T_max < 0 || V < 1) return fail(CORAL_EARG, "bad shape");
  if (B == 0) return CORAL_OK;
  if (!logits_dev || !out_tokens_dev || !out_lens_dev) return fail(CORAL_EARG, "coral_ctc_greedy: null buffer");
  cudaStream_t st = (cudaStream_t)stream;
  if (T_max == 0) {
    CORAL_CUDA_OK(cudaMemsetAsync(out_lens_dev, 0, sizeof(int32_t) * (size_t)B, st));
    return CORAL_OK;
  }
  // one stage = 128 rows + up to 12 bytes of lead-in, rounded to 128 bytes
  const size_t stage_bytes = (((size_t)kTileFrames * V * sizeof(float) + 16 + 127) / 128) * 128;
  const size_t smem = stage_bytes * kStages;
  if (smem > 200 * 1024) return fail(CORAL_EARG, "vocabulary too large for the greedy kernel's tiles");
  int dev = 0;
  CORAL_CUDA_OK(cudaGetDevice(&dev));
  // occupancy for this tile size, queried once per (device, shared-memory size)
  static int cache_per_sm[64] = {0};
  static size_t cache_smem[64] = {0};
  int per_sm = (cache_smem[dev & 63] == smem) ? cache_per_sm[dev & 63] : 0;
  if (per_sm == 0) {
    CORAL_CUDA_OK(cudaFuncSetAttribute(ctc_greedy_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CORAL_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ctc_greedy_fused_kernel, kTileFrames, smem));
    cache_per_sm[dev & 63] = per_sm;
    cache_smem[dev & 63] = smem;
  }
  int* work = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_greedy_mu);
    int*& w = g_greedy_work[std::make_pair(dev, (void*)st)];
    if (!w) {  // {claim counter, CTAs finished}: zeroed once, rewound by every launch's last CTA
      CORAL_CUDA_OK(cudaMalloc(&w, 2 * sizeof(int)));
      CORAL_CUDA_OK(cudaMemsetAsync(w, 0, 2 * sizeof(int), st));
    }
    work = w;
  }
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(B, (long long)std::max(per_sm, 1) * sm_count(dev)));
  ctc_greedy_fused_kernel<<<grid, kTileFrames, smem, st>>>(logits_dev, lengths_dev, B, T_max, V, blank_id, pad_fixup,
                                                            out_ids_dev, out_tokens_dev, out_lens_dev, work,
                                                            (unsigned)stage_bytes);
  CORAL_CUDA_OK(cudaGetLastError());
  return CORAL_OK;
}

}  // extern "C"
