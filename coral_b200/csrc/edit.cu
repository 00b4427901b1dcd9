// C-ABI: batched Levenshtein edit counts (substitutions, deletions, insertions, hits).
//
// Replaces jiwer.process_characters / process_words -> rapidfuzz Levenshtein.editops as
// called at R:src/coral/metrics.py:28 and :56 (UP:jiwer process.py/transforms.py,
// UP:rapidfuzz-cpp distance/Levenshtein_impl.hpp; behaviour per SURVEY.md section 8 A11/A12):
//   * jiwer's default transforms (cer: strip; wer: collapse whitespace runs, strip,
//     split on " ") are applied on the device to the raw code points;
//   * remove_common_affix: common prefix, then common suffix;
//   * edit_bitpar_kernel (pairs whose cores fit 64 * NW reference x 64 * NW hypothesis symbols, NW <= 4):
//     a warp takes 32 pairs; it prepares them one after the other with all lanes (transforms, affixes,
//     the pattern-match words of every hypothesis symbol through a small shared-memory hash), then
//     every lane runs Hyyro's bit-parallel recurrence -- 64 DP cells per word operation, the algorithm
//     rapidfuzz itself uses -- and the backtrace for ITS pair; the VP / VN words of every row go to an
//     interleaved per-warp area in HBM (coalesced across the lanes);
//   * edit_counts_kernel (everything longer): the unit-cost DP on anti-diagonals by one warp per pair
//     (lane = row of a 32-row strip), recording the same VP / VN bits (D[i][j] == D[i-1][j] +- 1);
//   * recover_alignment's backtrace preference (Delete, then Insert when the previous
//     row's VN bit is set, else the diagonal) gives the S / D / I split, which is NOT
//     determined by the distance alone (SURVEY 7 hard part 3).
// Integer work throughout; results are bit-exact against the oracle.
#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>
#include <utility>

#include "common.cuh"

namespace coral {

__device__ __forceinline__ bool is_space_cp(uint32_t c) {
  // str.isspace() / regex \s: 9-13, 28-32, 133, 160, 5760, 8192-8202, 8232, 8233, 8239, 8287, 12288
  if (c < 64u) return (0x1F0003E00ull >> c) & 1ull;
  if (c < 5760u) return c == 133u || c == 160u;
  return c == 5760u || c - 8192u <= 10u || c - 8232u <= 1u || c == 8239u || c == 8287u || c == 12288u;
}

// out_status value between the two passes of one call: "left for edit_counts_kernel"; never returned
constexpr int kEditDeferred = 3;

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
                   size_t gwork_stride, int gcap, int only_deferred) {
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
    if (only_deferred && out_status[pair] != kEditDeferred) continue;  // second pass: what the first left
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
    if (r1 < r0 || h1 < h0) { too_long = true; n1 = n2 = 0; }  // a span that ends before it begins: refused
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


// jiwer wer_default on a string of n <= 1024 symbols staged in shared memory, as 32-bit mask arithmetic:
// lane c owns positions 32c .. 32c+31. White-space bitmap -> strip -> separators (white space in a run
// of two or more, or a lone " ") -> word starts / ends. Same spans as split_words.
__device__ int split_words_bits(const uint32_t* raw, int n, int lane, uint32_t* wstart, uint32_t* wend, int cap) {
  constexpr unsigned full = 0xffffffffu;
  uint32_t W = 0, SP = 0;
  for (int c = 0; c * 32 < n; ++c) {
    const int i = c * 32 + lane;
    const uint32_t ch = i < n ? raw[i] : 0u;
    const unsigned mw = __ballot_sync(full, i < n && is_space_cp(ch));
    const unsigned ms = __ballot_sync(full, i < n && ch == 32u);
    if (lane == c) { W = mw; SP = ms; }
  }
  const int nch = (n + 31) >> 5;
  const uint32_t valid = lane < nch ? ((lane == nch - 1 && (n & 31)) ? ((1u << (n & 31)) - 1u) : full) : 0u;
  const uint32_t NS = ~W & valid;
  const unsigned anyns = __ballot_sync(full, NS != 0u);
  if (!anyns) return 0;
  const int cf = __ffs(anyns) - 1, cl = 31 - __clz(anyns);
  const int s = cf * 32 + __ffs(__shfl_sync(full, NS, cf)) - 1;
  const int e = cl * 32 + 32 - __clz(__shfl_sync(full, NS, cl));  // one past the last non-space symbol
  const int lo = lane * 32;
  uint32_t in = valid;
  if (s > lo) in &= (s - lo >= 32) ? 0u : (full << (s - lo));
  if (e < lo + 32) in &= (e <= lo) ? 0u : (full >> (lo + 32 - e));
  const uint32_t Wp = __shfl_up_sync(full, W, 1), Wn = __shfl_down_sync(full, W, 1);
  const uint32_t wl = (W << 1) | (lane ? Wp >> 31 : 0u);
  const uint32_t wr = (W >> 1) | (lane < 31 ? Wn << 31 : 0u);
  const uint32_t sep = W & (wl | wr | SP) & in;
  const uint32_t sp = __shfl_up_sync(full, sep, 1), sn = __shfl_down_sync(full, sep, 1);
  const uint32_t sl = (sep << 1) | (lane ? sp >> 31 : 0u);
  const uint32_t sr = (sep >> 1) | (lane < 31 ? sn << 31 : 0u);
  const uint32_t bs = (s >> 5) == lane ? 1u << (s & 31) : 0u;
  const uint32_t be = ((e - 1) >> 5) == lane ? 1u << ((e - 1) & 31) : 0u;
  const uint32_t body = in & ~sep;
  const uint32_t st = body & (sl | bs), en = body & (sr | be);
  int is = __popc(st), ie = __popc(en);  // inclusive scans over the lanes
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int a = __shfl_up_sync(full, is, d), b = __shfl_up_sync(full, ie, d);
    if (lane >= d) { is += a; ie += b; }
  }
  const int total = __shfl_sync(full, is, 31);
  int k = is - __popc(st);
  for (uint32_t m = st; m; m &= m - 1u, ++k)
    if (k < cap) wstart[k] = (uint32_t)(lo + __ffs(m) - 1);
  k = ie - __popc(en);
  for (uint32_t m = en; m; m &= m - 1u, ++k)
    if (k < cap) wend[k] = (uint32_t)(lo + __ffs(m));
  return total;
}

// ----------------------------------------------------------------------------- bit-parallel path
// Hyyro 2003 / rapidfuzz levenshtein_hyrroe2003_block (UP:rapidfuzz-cpp distance/Levenshtein_impl.hpp):
// the pattern (reference core, m1 symbols) lies along the bits, one row per hypothesis symbol.
// VP / VN of row j are exactly what recover_alignment reads (bit col-1 of row row-1).
//
// A warp takes a tile of 32 pairs.
//   A  (all lanes on one pair at a time) the raw strings arrive through a register prefetch one pair
//      ahead and are parked in shared memory, so that everything below runs without a global-memory
//      round trip: strip, (words: split, canonical word ids), common affixes, and the pattern-match
//      word PM[j][w] of every hypothesis symbol j through a 128-slot hash of pattern word w.
//   B  (lane = pair) the recurrence, 64 cells per word operation, PM words of four rows per load;
//      VP / VN of every row go to the warp's [row][word][lane] area in HBM (coalesced).
//   C  (lane = pair) the backtrace over a register window of eight rows per memory round trip.

constexpr uint32_t kEmptyKey = 0xffffffffu;
constexpr int kHashSlots = 128;  // >= 2 x the 64 symbols of one pattern word

__device__ __forceinline__ uint32_t edit_hash(uint32_t sym) { return (sym * 0x9E3779B1u) >> 25; }

__device__ __forceinline__ int64_t shfl_i64(int64_t v, int src) {
  const int lo = __shfl_sync(0xffffffffu, (int)(uint32_t)(uint64_t)v, src);
  const int hi = __shfl_sync(0xffffffffu, (int)(uint32_t)((uint64_t)v >> 32), src);
  return (int64_t)(((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo);
}

template <int NW, bool WORDS>
struct BitparShared {
  static constexpr int CAP = 64 * NW;                    // symbols (chars / tokens / words) per side
  static constexpr int RAW = WORDS ? 128 * NW : 64 * NW;  // raw code points per side
  uint32_t raw[2][RAW];
  uint32_t tok[WORDS ? 2 : 1][WORDS ? CAP : 1];   // word ids
  uint32_t span[WORDS ? 4 : 1][WORDS ? CAP : 1];  // word starts / (ends | hash16 << 16), ref then hyp
  uint32_t key[kHashSlots];
  uint32_t mask[kHashSlots * 2];
};

template <int NW, int MIN_CTAS, bool WORDS>
__global__ void __launch_bounds__(128, MIN_CTAS)
edit_bitpar_kernel(const uint32_t* __restrict__ ref_cps, const int64_t* __restrict__ ref_beg,
                   const int64_t* __restrict__ ref_end, const uint32_t* __restrict__ hyp_cps,
                   const int64_t* __restrict__ hyp_beg, const int64_t* __restrict__ hyp_end, int64_t n_pairs,
                   int mode, int32_t* __restrict__ out_sdih, int32_t* __restrict__ out_status,
                   ulonglong2* __restrict__ area, int defer_status, int ppt) {
  // ppt = pairs per tile (32, 16 or 8): a small batch is cut into more, smaller tiles so that every SM
  // has warps to run -- a tile's phase A is serial over its pairs, which is the kernel's latency
  constexpr int WARPS = 4;
  using Sh = BitparShared<NW, WORDS>;
  constexpr int CAP = Sh::CAP, RAW = Sh::RAW, KR = RAW / 32;
  __shared__ Sh s_all[WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Sh& sh = s_all[warp];
  const int64_t gw = (int64_t)blockIdx.x * WARPS + warp;
  const int64_t nwarps = (int64_t)gridDim.x * WARPS;
  // per warp: VP / VN rows [row][word][lane] (16 B per lane: coalesced in phase B), then the pattern-match
  // words [pair][word][row] (8 B: a pair's rows are consecutive, so phase A writes whole sectors)
  ulonglong2* rows = area + (size_t)gw * ((size_t)CAP * NW * 48);
  unsigned long long* pmw = reinterpret_cast<unsigned long long*>(rows + (size_t)CAP * NW * 32);
  uint32_t* key = sh.key;
  uint32_t* mask = sh.mask;
  for (int i = lane; i < kHashSlots; i += 32) { key[i] = kEmptyKey; mask[2 * i] = 0; mask[2 * i + 1] = 0; }
  __syncwarp();
  const int64_t n_tiles = (n_pairs + ppt - 1) / ppt;
  for (int64_t tile = gw; tile < n_tiles; tile += nwarps) {
    const int64_t base = tile * ppt;
    const int n_here = (int)min((int64_t)ppt, n_pairs - base);
    // this lane's pair
    int64_t my_r0 = 0, my_h0 = 0;
    int my_l1 = 0, my_l2 = 0;
    bool my_fits = false;
    if (lane < n_here) {
      my_r0 = ref_beg[base + lane];
      my_h0 = hyp_beg[base + lane];
      const int64_t l1 = ref_end[base + lane] - my_r0, l2 = hyp_end[base + lane] - my_h0;
      my_fits = l1 >= 0 && l2 >= 0 && l1 <= RAW && l2 <= RAW;
      my_l1 = my_fits ? (int)l1 : (l1 == 0 ? 0 : 1);  // only "empty or not" matters when it does not fit
      my_l2 = my_fits ? (int)l2 : 0;
    }
    int my_m1 = 0, my_m2 = 0, my_n1 = 0, my_status = 0;
    // ---- A
    uint32_t nx[2 * KR];
    auto fetch = [&](int q) {
      const int64_t r0 = shfl_i64(my_r0, q), h0 = shfl_i64(my_h0, q);
      const bool fits = __shfl_sync(0xffffffffu, (int)my_fits, q) != 0;
      const int l1 = fits ? __shfl_sync(0xffffffffu, my_l1, q) : 0;
      const int l2 = fits ? __shfl_sync(0xffffffffu, my_l2, q) : 0;
#pragma unroll
      for (int k = 0; k < KR; ++k) {
        const int i = lane + 32 * k;
        nx[k] = i < l1 ? ref_cps[r0 + i] : 0u;
        nx[KR + k] = i < l2 ? hyp_cps[h0 + i] : 0u;
      }
    };
    fetch(0);
    for (int q = 0; q < n_here; ++q) {
#pragma unroll
      for (int k = 0; k < KR; ++k) {
        sh.raw[0][lane + 32 * k] = nx[k];
        sh.raw[1][lane + 32 * k] = nx[KR + k];
      }
      __syncwarp();
      if (q + 1 < n_here) fetch(q + 1);  // in flight while this pair is prepared
      const bool fits = __shfl_sync(0xffffffffu, (int)my_fits, q) != 0;
      const int raw1 = __shfl_sync(0xffffffffu, my_l1, q);
      const int raw2 = __shfl_sync(0xffffffffu, my_l2, q);
      int status = raw1 == 0 ? 1 : 0;
      bool defer = !fits;
      const uint32_t* a = sh.raw[0];
      const uint32_t* b = sh.raw[1];
      int n1 = fits ? raw1 : 0, n2 = fits ? raw2 : 0;
      if (!defer && mode != CORAL_EDIT_TOKENS) {
        if (!WORDS) {
          int64_t rs, re, hs, he;
          strip_range(sh.raw[0], 0, raw1, lane, rs, re);
          strip_range(sh.raw[1], 0, raw2, lane, hs, he);
          a = sh.raw[0] + rs; n1 = (int)(re - rs);
          b = sh.raw[1] + hs; n2 = (int)(he - hs);
        }
        if (WORDS) {
          uint32_t* ws1 = sh.span[0]; uint32_t* we1 = sh.span[1];
          uint32_t* ws2 = sh.span[2]; uint32_t* we2 = sh.span[3];
          n1 = split_words_bits(sh.raw[0], raw1, lane, ws1, we1, CAP);
          n2 = split_words_bits(sh.raw[1], raw2, lane, ws2, we2, CAP);
          if (n1 > CAP || n2 > CAP) {
            defer = true;
            n1 = n2 = 0;
          } else {
            __syncwarp();
            // a 16-bit hash of every word rides in the upper half of its end offset ...
            for (int k = lane; k < n1 + n2; k += 32) {
              const bool kr = k < n1;
              const uint32_t* kc = kr ? sh.raw[0] : sh.raw[1];
              const uint32_t ks = kr ? ws1[k] : ws2[k - n1];
              const uint32_t ke = kr ? we1[k] : we2[k - n1];
              uint32_t h = 2166136261u;
              for (uint32_t t = ks; t < ke; ++t) h = (h ^ kc[t]) * 16777619u;
              h = (h ^ (h >> 16)) & 0xffffu;
              if (kr) we1[k] = ke | (h << 16); else we2[k - n1] = ke | (h << 16);
            }
            __syncwarp();
            // ... so that the canonical id of a word (index, in ref ++ hyp order, of its first exact
            // occurrence) compares code points only with words of the same length and hash
            for (int k = lane; k < n1 + n2; k += 32) {
              const bool kr = k < n1;
              const uint32_t* kc = kr ? sh.raw[0] : sh.raw[1];
              const uint32_t ks = kr ? ws1[k] : ws2[k - n1];
              const uint32_t kx = kr ? we1[k] : we2[k - n1];
              const uint32_t kl = (kx & 0xffffu) - ks;
              int id = k;
              for (int j = 0; j < k; ++j) {
                const bool jr = j < n1;
                const uint32_t jx = jr ? we1[j] : we2[j - n1];
                if ((jx ^ kx) >> 16) continue;
                const uint32_t js = jr ? ws1[j] : ws2[j - n1];
                if ((jx & 0xffffu) - js != kl) continue;
                const uint32_t* jc = jr ? sh.raw[0] : sh.raw[1];
                bool eq = true;
                for (uint32_t t = 0; t < kl; ++t)
                  if (kc[ks + t] != jc[js + t]) { eq = false; break; }
                if (eq) { id = j; break; }
              }
              if (kr) sh.tok[0][k] = (uint32_t)id; else sh.tok[WORDS ? 1 : 0][k - n1] = (uint32_t)id;
            }
            __syncwarp();
            a = sh.tok[0];
            b = sh.tok[WORDS ? 1 : 0];
          }
        }
      }
      if (!defer && n1 == 0) status = 1;
      // remove_common_affix
      int m1 = 0, m2 = 0, p = 0;
      if (!defer) {
        const int mn = min(n1, n2);
        for (int i0 = 0; i0 < mn; i0 += 32) {
          const int i = i0 + lane;
          const unsigned m = __ballot_sync(0xffffffffu, i < mn && a[i] != b[i]);
          if (m) { p = i0 + __ffs(m) - 1; break; }
          p = min(mn, i0 + 32);
        }
        int sfx = 0;
        const int mr = mn - p;
        for (int i0 = 0; i0 < mr; i0 += 32) {
          const int i = i0 + lane;
          const unsigned m = __ballot_sync(0xffffffffu, i < mr && a[n1 - 1 - i] != b[n2 - 1 - i]);
          if (m) { sfx = i0 + __ffs(m) - 1; break; }
          sfx = min(mr, i0 + 32);
        }
        m1 = n1 - p - sfx;
        m2 = n2 - p - sfx;
      }
      if (!defer && m1 > 0 && m2 > 0) {
        // PM[j][w] bit i = (a[p + 64 w + i] == b[p + j]), written where row j's VP will go
        const int nw = (m1 + 63) >> 6;
        bool bad = false;  // a token equal to the empty-slot marker (tokens mode only)
        for (int w = 0; w < nw; ++w) {
          int slot[2] = {-1, -1};
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int i = 64 * w + 32 * half + lane;
            if (i < m1) {
              const uint32_t sym = a[p + i];
              if (sym == kEmptyKey) {
                bad = true;
              } else {
                uint32_t h = edit_hash(sym);
                while (true) {
                  const uint32_t old = atomicCAS(&key[h], kEmptyKey, sym);
                  if (old == kEmptyKey || old == sym) break;
                  h = (h + 1) & (kHashSlots - 1);
                }
                atomicOr(&mask[2 * h + half], 1u << lane);
                slot[half] = (int)h;
              }
            }
          }
          __syncwarp();
          for (int j = lane; j < m2; j += 32) {
            const uint32_t sym = b[p + j];
            unsigned long long pm = 0;
            if (sym == kEmptyKey) {
              bad = true;
            } else {
              uint32_t h = edit_hash(sym);
              while (true) {
                const uint32_t k = key[h];
                if (k == sym) { pm = (unsigned long long)mask[2 * h] | ((unsigned long long)mask[2 * h + 1] << 32); break; }
                if (k == kEmptyKey) break;
                h = (h + 1) & (kHashSlots - 1);
              }
            }
            pmw[((size_t)q * NW + w) * CAP + j] = pm;
          }
          __syncwarp();
#pragma unroll
          for (int half = 0; half < 2; ++half)
            if (slot[half] >= 0) { key[slot[half]] = kEmptyKey; mask[2 * slot[half]] = 0; mask[2 * slot[half] + 1] = 0; }
          __syncwarp();
        }
        if (__any_sync(0xffffffffu, bad)) defer = true;
      }
      if (lane == q) {
        my_m1 = defer ? 0 : m1;
        my_m2 = defer ? 0 : m2;
        my_n1 = n1;
        my_status = defer ? defer_status : status;
      }
      __syncwarp();  // the next pair overwrites the staged strings
    }
    __syncwarp();
    // ---- B: every lane runs the recurrence for its own pair
    int D = my_m1, I = my_m2, S = 0;
    if (my_m1 > 0 && my_m2 > 0) {
      const int nw = (my_m1 + 63) >> 6;
      unsigned long long VP[NW], VN[NW];
#pragma unroll
      for (int w = 0; w < NW; ++w) { VP[w] = ~0ull; VN[w] = 0ull; }
      int dist = my_m1;
      const unsigned long long last = 1ull << ((my_m1 - 1) & 63);
      constexpr int PF = NW <= 2 ? 4 : 2;
      const unsigned long long* pml = pmw + (size_t)lane * NW * CAP;
      for (int j0 = 0; j0 < my_m2; j0 += PF) {
        unsigned long long pm[PF][NW];  // PF consecutive rows of a word: 16-byte loads (rows past m2 are unused)
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          if (w < nw) {
#pragma unroll
            for (int t = 0; t < PF; t += 2) {
              const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(pml + (size_t)w * CAP + j0 + t);
              pm[t][w] = v.x;
              pm[t + 1][w] = v.y;
            }
          } else {
#pragma unroll
            for (int t = 0; t < PF; ++t) pm[t][w] = 0ull;
          }
        }
#pragma unroll
        for (int t = 0; t < PF; ++t) {
          if (j0 + t < my_m2) {
            unsigned long long hp_c = 1, hn_c = 0;
            ulonglong2* r = rows + (size_t)(j0 + t) * NW * 32 + lane;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
              if (w < nw) {
                const unsigned long long X = pm[t][w] | hn_c;
                const unsigned long long D0 = (((X & VP[w]) + VP[w]) ^ VP[w]) | X | VN[w];
                const unsigned long long HP = VN[w] | ~(D0 | VP[w]);
                const unsigned long long HN = D0 & VP[w];
                if (w == nw - 1) dist += (int)((HP & last) != 0) - (int)((HN & last) != 0);
                const unsigned long long HPs = (HP << 1) | hp_c;
                const unsigned long long HNs = (HN << 1) | hn_c;
                hp_c = HP >> 63;
                hn_c = HN >> 63;
                VP[w] = HNs | ~(D0 | HPs);
                VN[w] = HPs & D0;
                r[w * 32] = make_ulonglong2(VP[w], VN[w]);
              }
            }
          }
        }
      }
      // ---- C: recover_alignment's walk (Delete, then Insert when the previous row's VN bit is set, else
      // the diagonal); substitutions = distance - deletions - insertions. Eight rows of the current
      // pattern word are fetched together; they serve seven row steps.
      constexpr int K = 8;
      int col = my_m1, row = my_m2;
      D = 0;
      I = 0;
      while (row > 0 && col > 0) {
        const int w = (col - 1) >> 6;
        const int top = row;
        ulonglong2 win[K];
#pragma unroll
        for (int t = 0; t < K; ++t) {
          const int r = top - 1 - t;
          win[t] = r >= 0 ? rows[((size_t)r * NW + w) * 32 + lane] : make_ulonglong2(0ull, 0ull);
        }
        bool go = true;
#pragma unroll
        for (int t = 0; t < K - 1; ++t) {
          if (go) {  // here row == top - t > 0, col > 0 and column col-1 lies in word w
            while (col > 0 && ((col - 1) >> 6) == w && ((win[t].x >> ((col - 1) & 63)) & 1ull)) { ++D; --col; }
            if (col == 0 || ((col - 1) >> 6) != w) {
              go = false;
            } else {
              --row;
              if (row > 0 && ((win[t + 1].y >> ((col - 1) & 63)) & 1ull)) ++I; else --col;
              if (row == 0 || col == 0 || ((col - 1) >> 6) != w) go = false;
            }
          }
        }
      }
      D += col;
      I += row;
      S = dist - D - I;
    }
    if (lane < n_here) {
      const bool refused = my_status >= 2;
      int32_t* o = out_sdih + (base + lane) * 4;
      o[0] = refused ? 0 : S;
      o[1] = refused ? 0 : D;
      o[2] = refused ? 0 : I;
      o[3] = refused ? 0 : my_n1 - (S + D);
      out_status[base + lane] = my_status;
    }
    __syncwarp();
  }
}

// Off-chip work areas: one per (device, stream, kernel), like the decoder's scratch arenas, so launches
// on different streams never share one. Grown lazily after waiting for that stream's earlier launches;
// kept for the life of the process.
struct EditArea {
  uint8_t* ptr = nullptr;
  size_t bytes = 0;
};
static std::mutex g_edit_mu;
static std::map<std::tuple<int, void*, int>, EditArea> g_edit_areas;

static int32_t edit_area(int device, cudaStream_t st, int which, size_t bytes, uint8_t** out) {
  std::lock_guard<std::mutex> lock(g_edit_mu);
  EditArea& A = g_edit_areas[std::make_tuple(device, (void*)st, which)];
  if (A.bytes < bytes) {
    CORAL_CUDA_OK(cudaStreamSynchronize(st));  // earlier launches on this stream may still use the old area
    if (A.ptr) cudaFree(A.ptr);
    A.ptr = nullptr;
    A.bytes = 0;
    CORAL_CUDA_OK(cudaMalloc(&A.ptr, bytes));
    A.bytes = bytes;
  }
  *out = A.ptr;
  return CORAL_OK;
}

template <int NW, int MIN_CTAS, bool WORDS>
static int32_t launch_bitpar(const uint32_t* ref_cps, const int64_t* ref_beg, const int64_t* ref_end,
                             const uint32_t* hyp_cps, const int64_t* hyp_beg, const int64_t* hyp_end,
                             int64_t n_pairs, int mode, int device, int32_t* out_sdih, int32_t* out_status,
                             int defer_status, cudaStream_t st) {
  const int sms = sm_count(device);
  const int64_t warps = (int64_t)sms * MIN_CTAS * 4;  // resident at once
  const int ppt = n_pairs >= 32 * warps ? 32 : (n_pairs >= 16 * warps ? 16 : 8);
  const int64_t tiles = (n_pairs + ppt - 1) / ppt;
  const unsigned grid = (unsigned)std::min<int64_t>((tiles + 3) / 4, (int64_t)sms * MIN_CTAS);
  const size_t per_warp = (size_t)(64 * NW) * NW * 48 * sizeof(ulonglong2);  // VP / VN rows + pattern-match words
  uint8_t* area = nullptr;
  const int32_t rc = edit_area(device, st, 1, per_warp * 4 * grid, &area);
  if (rc != CORAL_OK) return rc;
  edit_bitpar_kernel<NW, MIN_CTAS, WORDS><<<grid, 128, 0, st>>>(ref_cps, ref_beg, ref_end, hyp_cps, hyp_beg,
                                                                hyp_end, n_pairs, mode, out_sdih, out_status,
                                                                reinterpret_cast<ulonglong2*>(area), defer_status, ppt);
  return CORAL_OK;
}

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
  // First pass: the bit-parallel kernel, sized by the longest sequence the batch can hold (words: a word
  // and its separator take two symbols). What does not fit its 64 * NW cap -- and, in tokens mode, a
  // pair holding the token 0xffffffff -- is left for the general kernel in a second pass; when no second
  // pass follows, a pair that does not fit after all (a max_len that was too small) is refused (status 2).
  static const bool no_bitpar = std::getenv("CORAL_B200_EDIT_NO_BITPAR") != nullptr;
  const int64_t max_tokens = mode == CORAL_EDIT_WORDS ? (max_len + 1) / 2 : max_len;
  const bool second_pass = no_bitpar || max_tokens > 256 || mode == CORAL_EDIT_TOKENS;
  const int defer_status = second_pass ? kEditDeferred : 2;
  if (!no_bitpar) {
    int32_t rc;
#define CORAL_BITPAR(NW, CTAS, CTASW)                                                                                 \
  rc = mode == CORAL_EDIT_WORDS                                                                                  \
           ? launch_bitpar<NW, CTASW, true>(ref_cps_dev, ref_begin_dev, ref_end_dev, hyp_cps_dev, hyp_begin_dev,  \
                                           hyp_end_dev, n_pairs, mode, device, out_sdih_dev, out_status_dev,     \
                                           defer_status, st)                                                     \
           : launch_bitpar<NW, CTAS, false>(ref_cps_dev, ref_begin_dev, ref_end_dev, hyp_cps_dev, hyp_begin_dev, \
                                            hyp_end_dev, n_pairs, mode, device, out_sdih_dev, out_status_dev,    \
                                            defer_status, st)
    if (max_tokens <= 64) CORAL_BITPAR(1, 6, 6);
    else if (max_tokens <= 128) CORAL_BITPAR(2, 4, 5);
    else if (max_tokens <= 192) CORAL_BITPAR(3, 3, 3);
    else CORAL_BITPAR(4, 2, 2);
#undef CORAL_BITPAR
    if (rc != CORAL_OK) return rc;
    CORAL_CUDA_OK(cudaGetLastError());
  }
  if (second_pass) {
    constexpr int WARPS = 8;
    const int only_deferred = no_bitpar ? 0 : 1;
    const int64_t need = (n_pairs + WARPS - 1) / WARPS;
    if (max_len <= 128) {
      const unsigned grid = (unsigned)std::min<int64_t>(need, (int64_t)sms * 4);
      edit_counts_kernel<false, 128, WARPS><<<grid, WARPS * 32, 0, st>>>(
          ref_cps_dev, ref_begin_dev, ref_end_dev, hyp_cps_dev, hyp_begin_dev, hyp_end_dev, n_pairs, mode,
          out_sdih_dev, out_status_dev, nullptr, 0, 0, only_deferred);
    } else {
      const int cap = (int)std::max<int64_t>(64, (max_len + 31) / 32 * 32);
      const size_t stride = ((size_t)cap * 4 * 2 + ((size_t)cap + 32) * 4 + (size_t)cap * (cap / 32) * 4 * 2 + 15) & ~(size_t)15;
      // long strings: fewer resident warps keep the work area bounded (1 MiB per warp at 2048 symbols)
      const unsigned grid = (unsigned)std::min<int64_t>(need, (int64_t)sms * (cap <= 512 ? 4 : 1));
      uint8_t* work = nullptr;
      const int32_t rc = edit_area(device, st, 0, stride * WARPS * grid, &work);
      if (rc != CORAL_OK) return rc;
      edit_counts_kernel<true, 128, WARPS><<<grid, WARPS * 32, 0, st>>>(
          ref_cps_dev, ref_begin_dev, ref_end_dev, hyp_cps_dev, hyp_begin_dev, hyp_end_dev, n_pairs, mode,
          out_sdih_dev, out_status_dev, work, stride, cap, only_deferred);
    }
    CORAL_CUDA_OK(cudaGetLastError());
  }
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
