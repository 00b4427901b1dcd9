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
// utterance) and writes T*4 bytes; nothing is re-read. A persistent grid streams tiles of
// 128 frames through a two-stage cp.async pipeline in shared memory.
#include <algorithm>

#include "common.cuh"

namespace coral {

constexpr int kTileFrames = 128;  // frames per pipeline stage (= threads per CTA)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}

// Asynchronous copy of one tile (nfr * V contiguous floats) into a shared-memory stage, at
// the widest granularity the source alignment allows (rows are 184 B: 8-byte aligned always,
// 16-byte aligned for every other frame).
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
}

// Persistent, double-buffered: while the CTA scans the rows of one tile, the next tile is in
// flight (cp.async, no registers involved). Tiles are (utterance, 128-frame block); blocks
// past an utterance's length are skipped. Thread f scans row f of the tile from shared memory.
__global__ void __launch_bounds__(kTileFrames)
ctc_argmax_kernel(const float* __restrict__ logits, const int32_t* __restrict__ lengths, int B, int T_max, int V,
                  int blank_id, int pad_fixup, int32_t* __restrict__ out_ids) {
  extern __shared__ __align__(16) float smem[];
  const int stage_floats = (kTileFrames * V + 3) & ~3;
  const int tpu = (T_max + kTileFrames - 1) / kTileFrames;
  const long long n_tiles = (long long)B * tpu;
  auto tile_frames = [&](long long tile, int& u, int& t0) -> int {
    u = (int)(tile / tpu);
    t0 = (int)(tile % tpu) * kTileFrames;
    const int T = lengths ? lengths[u] : T_max;
    return T - t0 < kTileFrames ? T - t0 : kTileFrames;  // <= 0: nothing to do
  };
  auto next_valid = [&](long long tile) -> long long {
    int u, t0;
    while (tile < n_tiles && tile_frames(tile, u, t0) <= 0) tile += gridDim.x;
    return tile;
  };
  long long cur = next_valid(blockIdx.x);
  int st = 0;
  if (cur < n_tiles) {
    int u, t0;
    const int nfr = tile_frames(cur, u, t0);
    issue_tile(logits + ((size_t)u * T_max + t0) * V, smem, nfr * V);
  }
  while (cur < n_tiles) {
    const long long nxt = next_valid(cur + gridDim.x);
    if (nxt < n_tiles) {
      int u, t0;
      const int nfr = tile_frames(nxt, u, t0);
      issue_tile(logits + ((size_t)u * T_max + t0) * V, smem + (st ^ 1) * stage_floats, nfr * V);
      asm volatile("cp.async.wait_group 1;");
    } else {
      asm volatile("cp.async.wait_group 0;");
    }
    __syncthreads();
    int u, t0;
    const int nfr = tile_frames(cur, u, t0);
    const int f = threadIdx.x;
    if (f < nfr) {
      const float* row = smem + st * stage_floats + f * V;
      float best = row[0];
      int id = 0;
      bool all_m100 = best == -100.0f;
      for (int v = 1; v < V; ++v) {
        const float x = row[v];
        all_m100 &= x == -100.0f;
        // first maximum wins; like numpy, the first NaN wins over everything
        if (x > best || (x != x && best == best)) { best = x; id = v; }
      }
      if (pad_fixup && all_m100) id = blank_id;
      out_ids[(size_t)u * T_max + t0 + f] = id;
    }
    __syncthreads();  // the stage is free again before the next iteration refills it
    cur = nxt;
    st ^= 1;
  }
}

// One CTA per utterance: keep[t] = id != blank && (!group || t == 0 || id != ids[t-1]),
// compacted with a block scan. Safe in place (writes never pass the read cursor).
__global__ void __launch_bounds__(256)
ctc_collapse_kernel(const int32_t* ids, const int32_t* __restrict__ lengths, int T_max, int blank_id,
                    int group_tokens, int32_t* out_tokens, int32_t* __restrict__ out_lens) {
  __shared__ int warp_tot[8];
  __shared__ int base_s;
  const int u = blockIdx.x;
  const int T = lengths ? lengths[u] : T_max;
  const int32_t* src = ids + (size_t)u * T_max;
  int32_t* dst = out_tokens + (size_t)u * T_max;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (int c0 = 0; c0 < T; c0 += 256) {
    const int t = c0 + threadIdx.x;
    int id = 0, keep = 0;
    if (t < T) {
      id = src[t];
      const int prev = (t > 0) ? src[t - 1] : -1;
      keep = (id != blank_id) && (!group_tokens || t == 0 || id != prev);
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    const int in_warp = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[warp] = __popc(m);
    __syncthreads();  // all reads of this chunk are done, warp totals visible
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += warp_tot[w];
    const int base = base_s;
    if (keep) dst[base + woff + in_warp] = id;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < 8; ++w) tot += warp_tot[w];
      base_s = base + tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out_lens[u] = base_s;
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
  ctc_collapse_kernel<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(ids_dev, lengths_dev, T_max, blank_id,
                                                                    group_tokens, out_tokens_dev, out_lens_dev);
  CORAL_CUDA_OK(cudaGetLastError());
  return CORAL_OK;
}

int32_t coral_ctc_greedy(const float* logits_dev, const int32_t* lengths_dev, int32_t B, int32_t T_max, int32_t V,
                         int32_t blank_id, int32_t pad_fixup, int32_t* out_ids_dev, int32_t* out_tokens_dev,
                         int32_t* out_lens_dev, void* stream) {
  if (B < 0 || T_max < 0 || V < 1) return fail(CORAL_EARG, "bad shape");
  if (B == 0) return CORAL_OK;
  if (!logits_dev || !out_tokens_dev || !out_lens_dev) return fail(CORAL_EARG, "coral_ctc_greedy: null buffer");
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* ids = out_ids_dev ? out_ids_dev : out_tokens_dev;
  if (T_max > 0) {
    const size_t stage_floats = ((size_t)kTileFrames * V + 3) & ~(size_t)3;
    const size_t smem = 2 * stage_floats * sizeof(float);
    if (smem > 200 * 1024) return fail(CORAL_EARG, "vocabulary too large for the greedy kernel's tiles");
    CORAL_CUDA_OK(cudaFuncSetAttribute(ctc_argmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, per_sm = 0;
    CORAL_CUDA_OK(cudaGetDevice(&dev));
    CORAL_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ctc_argmax_kernel, kTileFrames, smem));
    const long long n_tiles = (long long)B * ((T_max + kTileFrames - 1) / kTileFrames);
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(n_tiles, (long long)std::max(per_sm, 1) * sm_count(dev)));
    ctc_argmax_kernel<<<grid, kTileFrames, smem, st>>>(logits_dev, lengths_dev, B, T_max, V, blank_id, pad_fixup, ids);
    CORAL_CUDA_OK(cudaGetLastError());
  }
  return coral_ctc_collapse(ids, lengths_dev, B, T_max, blank_id, 1, out_tokens_dev, out_lens_dev, stream);
}

}  // extern "C"
