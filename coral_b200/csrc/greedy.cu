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
// utterance) and writes T*4 bytes; nothing is re-read. Tiles of 256 frames are staged
// in shared memory with 16-byte coalesced loads, rows padded to an odd stride so the
// per-thread row scan is bank-conflict free.
#include "common.cuh"

namespace coral {

constexpr int kFramesPerTile = 256;

template <typename VecT>
__device__ __forceinline__ void stage_tile(const float* __restrict__ src, float* __restrict__ tile, int n_floats,
                                           int V, int VP) {
  constexpr int W = sizeof(VecT) / 4;
  const int nvec = n_floats / W;
  const VecT* s = reinterpret_cast<const VecT*>(src);
#pragma unroll 4
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    const VecT v = __ldcs(s + i);  // streamed once: evict-first
    const float* f = reinterpret_cast<const float*>(&v);
#pragma unroll
    for (int k = 0; k < W; ++k) {
      const int e = i * W + k;
      tile[(e / V) * VP + (e % V)] = f[k];
    }
  }
  for (int e = nvec * W + threadIdx.x; e < n_floats; e += blockDim.x) tile[(e / V) * VP + (e % V)] = src[e];
}

__global__ void __launch_bounds__(kFramesPerTile)
ctc_argmax_kernel(const float* __restrict__ logits, const int32_t* __restrict__ lengths, int T_max, int V, int VP,
                  int blank_id, int pad_fixup, int32_t* __restrict__ out_ids) {
  extern __shared__ float tile[];
  const int u = blockIdx.y;
  const int t0 = blockIdx.x * kFramesPerTile;
  const int T = lengths ? lengths[u] : T_max;
  if (t0 >= T) return;
  const int nfr = min(kFramesPerTile, T - t0);
  const float* src = logits + ((size_t)u * T_max + t0) * V;
  const int n = nfr * V;
  const uintptr_t a = reinterpret_cast<uintptr_t>(src);
  if ((a & 15) == 0) stage_tile<float4>(src, tile, n, V, VP);
  else if ((a & 7) == 0) stage_tile<float2>(src, tile, n, V, VP);
  else stage_tile<float>(src, tile, n, V, VP);
  __syncthreads();
  const int f = threadIdx.x;
  if (f < nfr) {
    const float* row = tile + f * VP;
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
  if (B > 65535) return fail(CORAL_EARG, "batch above 65535: split the call");
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* ids = out_ids_dev ? out_ids_dev : out_tokens_dev;
  if (T_max > 0) {
    const int VP = V | 1;
    const size_t smem = (size_t)kFramesPerTile * VP * sizeof(float);
    if (smem > 48 * 1024)
      CORAL_CUDA_OK(cudaFuncSetAttribute(ctc_argmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((T_max + kFramesPerTile - 1) / kFramesPerTile), (unsigned)B);
    ctc_argmax_kernel<<<grid, kFramesPerTile, smem, st>>>(logits_dev, lengths_dev, T_max, V, VP, blank_id, pad_fixup,
                                                        ids);
    CORAL_CUDA_OK(cudaGetLastError());
  }
  return coral_ctc_collapse(ids, lengths_dev, B, T_max, blank_id, 1, out_tokens_dev, out_lens_dev, stream);
}

}  // extern "C"
