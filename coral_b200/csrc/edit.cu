// C-ABI: batched Levenshtein edit counts (substitutions, deletions, insertions, hits).
//
// Replaces jiwer.process_characters / process_words -> rapidfuzz Levenshtein.editops as
// called at R:src/coral/metrics.py:28 and :56 (UP:jiwer process.py/transforms.py,
// UP:rapidfuzz-cpp distance/Levenshtein_impl.hpp; behaviour per SURVEY.md section 8 A11/A12):
//   * jiwer's default transforms (cer: strip; wer: collapse whitespace runs, strip,
//     split on " ") are applied on the device to the raw code points;
//   * remove_common_affix: common prefix, then common suffix;
//   * the unit-cost DP is evaluated on anti-diagonals by one warp per pair (lane = row of
//     a 32-row strip) and records rapidfuzz's VP / VN bits (D[i][j] == D[i-1][j] +- 1);
//   * recover_alignment's backtrace preference (Delete, then Insert when the previous
//     row's VN bit is set, else the diagonal) gives the S / D / I split, which is NOT
//     determined by the distance alone (SURVEY 7 hard part 3).
// Integer work throughout; results are bit-exact against the oracle.
#include <algorithm>
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace coral {

__device__ __forceinline__ bool is_space_cp(uint32_t c) {
  // str.isspace() / regex \s for the code points that can occur
  return (c >= 9 && c <= 13) || (c >= 28 && c <= 32) || c == 133 || c == 160 || c == 5760 ||
         (c >= 8192 && c <= 8202) || c == 8232 || c == 8233 || c == 8239 || c == 8287 || c == 12288;
}

struct EditWork {
  uint32_t* tok1;
  uint32_t* tok2;
  int32_t* edge;
  uint32_t* vp;
  uint32_t* vn;
  int cap;  // tokens per side (multiple of 32)
};

// strip(): [s, e) of the non-whitespace core
__device__ void strip_range(const uint32_t* cps, int64_t a0, int64_t a1, int lane, int64_t& s, int64_t& e) {
  s = a1;
  for (int64_t i0 = a0; i0 < a1; i0 += 32) {
    const int64_t i = i0 + lane;
    const unsigned m = __ballot_sync(0xffffffffu, i < a1 && !is_space_cp(cps[i]));
    if (m) { s = i0 + (__ffs(m) - 1); break; }
  }
  e = s;
  if (s == a1) return;
  for (int64_t i1 = a1; i1 > s; i1 -= 32) {
    const int64_t i = i1 - 1 - lane;
    const unsigned m = __ballot_sync(0xffffffffu, i >= s && !is_space_cp(cps[i]));
    if (m) { e = i1 - (__ffs(m) - 1); break; }
  }
}

// jiwer wer_default on [s, e): word k spans [wstart[k], wend[k]). Returns the word count.
__device__ int split_words(const uint32_t* cps, int64_t s, int64_t e, int lane, uint32_t* wstart, uint32_t* wend,
                           int64_t base_off, int cap) {
  int nw = 0;
  for (int64_t i0 = s; i0 < e; i0 += 32) {
    const int64_t i = i0 + lane;
    bool st = false, en = false;
    if (i < e) {
      const uint32_t c = cps[i];
      const bool ws = is_space_cp(c);
      const bool wsl = i > s && is_space_cp(cps[i - 1]);
      const bool wsr = i + 1 < e && is_space_cp(cps[i + 1]);
      const bool sep = ws && (wsl || wsr || c == 32u);
      if (!sep) {
        // neighbours are separators iff they are whitespace in a run of >= 2 or a lone " "
        bool sepl = false, sepr = false;
        if (i > s) {
          const uint32_t cl = cps[i - 1];
          const bool wsll = i - 1 > s && is_space_cp(cps[i - 2]);
          sepl = wsl && (wsll || ws || cl == 32u);
        }
        if (i + 1 < e) {
          const uint32_t cr = cps[i + 1];
          const bool wsrr = i + 2 < e && is_space_cp(cps[i + 2]);
          sepr = wsr && (ws || wsrr || cr == 32u);
        }
        st = (i == s) || sepl;
        en = (i + 1 == e) || sepr;
      }
    }
    const unsigned ms = __ballot_sync(0xffffffffu, st);
    const unsigned le = lane == 31 ? 0xffffffffu : ((1u << (lane + 1)) - 1u);
    const int k = nw + __popc(ms & le) - 1;  // index of the word this position belongs to
    if (st && k < cap) wstart[k] = (uint32_t)(i - base_off);
    if (en && k < cap) wend[k] = (uint32_t)(i + 1 - base_off);
    nw += __popc(ms);
  }
  return nw;
}

template <bool GLOBAL_WORK, int LCAP, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
edit_counts_kernel(const uint32_t* __restrict__ ref_cps, const int64_t* __restrict__ ref_beg,
                   const int64_t* __restrict__ ref_end, const uint32_t* __restrict__ hyp_cps,
                   const int64_t* __restrict__ hyp_beg, const int64_t* __restrict__ hyp_end, int64_t n_pairs,
                   int mode, int32_t* __restrict__ out_sdih, int32_t* __restrict__ out_status, uint8_t* gwork,
                   size_t gwork_stride, int gcap) {
  // Every warp owns a shared-memory work area for strings of up to LCAP symbols; with
  // GLOBAL_WORK it also owns a larger one in HBM and picks per pair: only the pairs that do
  // not fit on chip pay for the off-chip matrix.
  constexpr int SCAP = LCAP;
  __shared__ uint32_t s_tok[WARPS][2][SCAP];
  __shared__ int32_t s_edge[WARPS][SCAP + 1];
  __shared__ uint32_t s_bits[WARPS][2][SCAP * (SCAP / 32)];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t gw = (int64_t)blockIdx.x * WARPS + warp;
  const int64_t nwarps = (int64_t)gridDim.x * WARPS;
  EditWork Wg, Ws;
  if (GLOBAL_WORK) {
    uint8_t* b = gwork + (size_t)gw * gwork_stride;
    Wg.cap = gcap;
    Wg.tok1 = reinterpret_cast<uint32_t*>(b); b += (size_t)gcap * 4;
    Wg.tok2 = reinterpret_cast<uint32_t*>(b); b += (size_t)gcap * 4;
    Wg.edge = reinterpret_cast<int32_t*>(b); b += ((size_t)gcap + 32) * 4;
    Wg.vp = reinterpret_cast<uint32_t*>(b); b += (size_t)gcap * (gcap / 32) * 4;
    Wg.vn = reinterpret_cast<uint32_t*>(b);
  }
  Ws.cap = LCAP;
  Ws.tok1 = s_tok[warp][0];
  Ws.tok2 = s_tok[warp][1];
  Ws.edge = s_edge[warp];
  Ws.vp = s_bits[warp][0];
  Ws.vn = s_bits[warp][1];
  for (int64_t pair = gw; pair < n_pairs; pair += nwarps) {
    __syncwarp();
    const int64_t r0 = ref_beg[pair], r1 = ref_end[pair];
    const int64_t h0 = hyp_beg[pair], h1 = hyp_end[pair];
    const bool on_chip = !GLOBAL_WORK || (r1 - r0 <= LCAP && h1 - h0 <= LCAP);
    const EditWork W = on_chip ? Ws : Wg;
    int n1 = 0, n2 = 0;
    int status = (r1 == r0) ? 1 : 0;
    bool too_long = false;  // a string that does not fit the work area is reported, never truncated
    if (mode == CORAL_EDIT_TOKENS) {
      too_long = r1 - r0 > W.cap || h1 - h0 > W.cap;
      n1 = too_long ? 0 : (int)(r1 - r0);
      n2 = too_long ? 0 : (int)(h1 - h0);
      for (int i = lane; i < n1; i += 32) W.tok1[i] = ref_cps[r0 + i];
      for (int i = lane; i < n2; i += 32) W.tok2[i] = hyp_cps[h0 + i];
    } else {
      int64_t rs, re, hs, he;
      strip_range(ref_cps, r0, r1, lane, rs, re);
      strip_range(hyp_cps, h0, h1, lane, hs, he);
      if (mode == CORAL_EDIT_CHARS) {
        too_long = re - rs > W.cap || he - hs > W.cap;
        n1 = too_long ? 0 : (int)(re - rs);
        n2 = too_long ? 0 : (int)(he - hs);
        for (int i = lane; i < n1; i += 32) W.tok1[i] = ref_cps[rs + i];
        for (int i = lane; i < n2; i += 32) W.tok2[i] = hyp_cps[hs + i];
      } else {
        // words: spans go to the (not yet used) bit-matrix area, ids to tok1/tok2
        uint32_t* ws1 = W.vp;              // [cap] starts (ref), relative to r0
        uint32_t* we1 = W.vp + W.cap;      // needs cap*(cap/32) >= 2*cap  <=> cap >= 64
        uint32_t* ws2 = W.vn;
        uint32_t* we2 = W.vn + W.cap;
        n1 = split_words(ref_cps, rs, re, lane, ws1, we1, r0, W.cap);
        n2 = split_words(hyp_cps, hs, he, lane, ws2, we2, h0, W.cap);
        if (n1 > W.cap || n2 > W.cap) { too_long = true; n1 = n2 = 0; }
        __syncwarp();
        // canonical id of a word = index (in ref ++ hyp order) of its first exact occurrence
        for (int k = lane; k < n1 + n2; k += 32) {
          const bool kr = k < n1;
          const uint32_t* kc = kr ? ref_cps + r0 : hyp_cps + h0;
          const uint32_t ks = kr ? ws1[k] : ws2[k - n1];
          const uint32_t kl = (kr ? we1[k] : we2[k - n1]) - ks;
          int id = k;
          for (int j = 0; j < k; ++j) {
            const bool jr = j < n1;
            const uint32_t* jc = jr ? ref_cps + r0 : hyp_cps + h0;
            const uint32_t js = jr ? ws1[j] : ws2[j - n1];
            const uint32_t jl = (jr ? we1[j] : we2[j - n1]) - js;
            if (jl != kl) continue;
            bool eq = true;
            for (uint32_t q = 0; q < kl; ++q)
              if (kc[ks + q] != jc[js + q]) { eq = false; break; }
            if (eq) { id = j; break; }
          }
          if (kr) W.tok1[k] = (uint32_t)id; else W.tok2[k - n1] = (uint32_t)id;
        }
      }
    }
    if (n1 == 0) status = 1;
    if (too_long) status = 2;
    __syncwarp();
    // remove_common_affix
    int p = 0;
    {
      const int mn = min(n1, n2);
      for (int i0 = 0; i0 < mn; i0 += 32) {
        const int i = i0 + lane;
        const unsigned m = __ballot_sync(0xffffffffu, i < mn && W.tok1[i] != W.tok2[i]);
        if (m) { p = i0 + __ffs(m) - 1; break; }
        p = min(mn, i0 + 32);
      }
    }
    int sfx = 0;
    {
      const int mn = min(n1, n2) - p;
      for (int i0 = 0; i0 < mn; i0 += 32) {
        const int i = i0 + lane;
        const unsigned m = __ballot_sync(0xffffffffu, i < mn && W.tok1[n1 - 1 - i] != W.tok2[n2 - 1 - i]);
        if (m) { sfx = i0 + __ffs(m) - 1; break; }
        sfx = min(mn, i0 + 32);
      }
    }
    const int m1 = n1 - p - sfx, m2 = n2 - p - sfx;
    const uint32_t* a = W.tok1 + p;
    const uint32_t* b = W.tok2 + p;
    int S = 0, D = 0, I = 0;
    // rapidfuzz aligns directly only while 2 * len1 * len2 bits stay under 1 MiB (then it splits the
    // problem Hirschberg-style, which may pick another optimal script): beyond that, refuse
    if ((int64_t)m1 * m2 >= 4194304 && m1 >= 65 && m2 >= 10) status = 2;
    if (status == 2) {
      // reported through out_status; the counts are zeroed
    } else if (m1 == 0 || m2 == 0) {
      D = m1;
      I = m2;
    } else {
      const int wpr = (m2 + 31) / 32;
      for (int j = lane; j <= m2; j += 32) W.edge[j] = j;  // D[0][j]
      __syncwarp();
      for (int i0 = 0; i0 < m1; i0 += 32) {
        const int i = i0 + lane + 1;
        const bool row_ok = i <= m1;
        const uint32_t mytok = row_ok ? a[i - 1] : 0u;
        int d = i;             // D[i][0]
        int up_prev = i - 1;   // D[i-1][0]
        uint32_t vpw = 0, vnw = 0;
        const int steps = m2 + 31;
        for (int k = 0; k < steps; ++k) {
          const int j = k - lane + 1;
          const int t = __shfl_up_sync(0xffffffffu, d, 1);
          if (j >= 1 && j <= m2) {
            const int up = lane == 0 ? W.edge[j] : t;
            const int cost = mytok != b[j - 1];
            int dn = min(min(up + 1, d + 1), up_prev + cost);
            const int bit = (j - 1) & 31;
            vpw |= (uint32_t)(dn == up + 1) << bit;
            vnw |= (uint32_t)(dn == up - 1) << bit;
            if (row_ok && (bit == 31 || j == m2)) {
              W.vp[(size_t)(i - 1) * wpr + ((j - 1) >> 5)] = vpw;
              W.vn[(size_t)(i - 1) * wpr + ((j - 1) >> 5)] = vnw;
            }
            if (bit == 31) { vpw = 0; vnw = 0; }
            up_prev = up;
            d = dn;
            if (lane == 31) W.edge[j] = dn;  // row i0+32 feeds the next strip
          }
        }
        __syncwarp();
        if (lane == 0) W.edge[0] = i0 + 32;
        __syncwarp();
      }
      if (lane == 0) {
        int col = m1, row = m2;
        while (row && col) {
          const uint32_t vpb = (W.vp[(size_t)(col - 1) * wpr + ((row - 1) >> 5)] >> ((row - 1) & 31)) & 1u;
          if (vpb) {
            ++D;
            --col;
          } else {
            --row;
            if (row && ((W.vn[(size_t)(col - 1) * wpr + ((row - 1) >> 5)] >> ((row - 1) & 31)) & 1u)) {
              ++I;
            } else {
              --col;
              if (a[col] != b[row]) ++S;
            }
          }
        }
        D += col;
        I += row;
      }
    }
    if (lane == 0) {
      int32_t* o = out_sdih + pair * 4;
      o[0] = S;
      o[1] = D;
      o[2] = I;
      o[3] = status == 2 ? 0 : n1 - (S + D);
      out_status[pair] = status;
    }
  }
}

// Off-chip DP work areas for pairs longer than the on-chip buffers: one per (device, stream),
// like the decoder's scratch arenas, so launches on different streams never share one. Grown
// lazily after waiting for that stream's earlier launches; kept for the life of the process.
struct EditArea {
  uint8_t* ptr = nullptr;
  size_t bytes = 0;
};
static std::mutex g_edit_mu;
static std::map<std::pair<int, void*>, EditArea> g_edit_areas;

}  // namespace coral

using namespace coral;

extern "C" {

int32_t coral_edit_counts_spans(const uint32_t* ref_cps_dev, const int64_t* ref_begin_dev,
                                const int64_t* ref_end_dev, const uint32_t* hyp_cps_dev,
                                const int64_t* hyp_begin_dev, const int64_t* hyp_end_dev, int64_t n_pairs,
                                int32_t mode, int64_t max_len, int32_t device, int32_t* out_sdih_dev,
                                int32_t* out_status_dev, void* stream) {
  if (n_pairs < 0 || max_len < 0) return fail(CORAL_EARG, "negative size");
  if (mode < 0 || mode > 2) return fail(CORAL_EARG, "mode must be 0 (tokens), 1 (chars) or 2 (words)");
  if (n_pairs == 0) return CORAL_OK;
  if (!ref_begin_dev || !ref_end_dev || !hyp_begin_dev || !hyp_end_dev || !out_sdih_dev || !out_status_dev)
    return fail(CORAL_EARG, "coral_edit_counts: null buffer");
  if (device < 0) return fail(CORAL_EARG, "device index out of range");
  if (max_len > 2048)
    return fail(CORAL_ECAP, "strings above 2048 symbols are outside rapidfuzz's direct-alignment range "
                            "(it switches to Hirschberg splitting there): not restated");
  DeviceGuard g(device);
  cudaStream_t st = (cudaStream_t)stream;
  const int sms = sm_count(device);
  if (max_len <= 128) {
    constexpr int WARPS = 8;
    const int64_t need = (n_pairs + WARPS - 1) / WARPS;
    const unsigned grid = (unsigned)std::min<int64_t>(need, (int64_t)sms * 4);
    edit_counts_kernel<false, 128, WARPS><<<grid, WARPS * 32, 0, st>>>(
        ref_cps_dev, ref_begin_dev, ref_end_dev, hyp_cps_dev, hyp_begin_dev, hyp_end_dev, n_pairs, mode,
        out_sdih_dev, out_status_dev, nullptr, 0, 0);
  } else {
    constexpr int WARPS = 8;
    const int cap = (int)std::max<int64_t>(64, (max_len + 31) / 32 * 32);
    const size_t stride = ((size_t)cap * 4 * 2 + ((size_t)cap + 32) * 4 + (size_t)cap * (cap / 32) * 4 * 2 + 15) & ~(size_t)15;
    const int64_t need = (n_pairs + WARPS - 1) / WARPS;
    // long strings: fewer resident warps keep the work area bounded (1 MiB per warp at 2048 symbols)
    const unsigned grid = (unsigned)std::min<int64_t>(need, (int64_t)sms * (cap <= 512 ? 4 : 1));
    const size_t bytes = stride * WARPS * grid;
    uint8_t* work = nullptr;
    {
      std::lock_guard<std::mutex> lock(g_edit_mu);
      EditArea& A = g_edit_areas[std::make_pair(device, (void*)st)];
      if (A.bytes < bytes) {
        CORAL_CUDA_OK(cudaStreamSynchronize(st));  // earlier launches on this stream may still use the old area
        if (A.ptr) cudaFree(A.ptr);
        A.ptr = nullptr;
        A.bytes = 0;
        CORAL_CUDA_OK(cudaMalloc(&A.ptr, bytes));
        A.bytes = bytes;
      }
      work = A.ptr;
    }
    edit_counts_kernel<true, 128, WARPS><<<grid, WARPS * 32, 0, st>>>(
        ref_cps_dev, ref_begin_dev, ref_end_dev, hyp_cps_dev, hyp_begin_dev, hyp_end_dev, n_pairs, mode,
        out_sdih_dev, out_status_dev, work, stride, cap);
  }
  CORAL_CUDA_OK(cudaGetLastError());
  return CORAL_OK;
}

int32_t coral_edit_counts(const uint32_t* ref_cps_dev, const int64_t* ref_offsets_dev, const uint32_t* hyp_cps_dev,
                          const int64_t* hyp_offsets_dev, int64_t n_pairs, int32_t mode, int64_t max_len,
                          int32_t device, int32_t* out_sdih_dev, int32_t* out_status_dev, void* stream) {
  if (!ref_offsets_dev || !hyp_offsets_dev) return fail(CORAL_EARG, "coral_edit_counts: null buffer");
  return coral_edit_counts_spans(ref_cps_dev, ref_offsets_dev, ref_offsets_dev + 1, hyp_cps_dev, hyp_offsets_dev,
                                 hyp_offsets_dev + 1, n_pairs, mode, max_len, device, out_sdih_dev, out_status_dev,
                                 stream);
}

}  // extern "C"
