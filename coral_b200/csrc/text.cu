// C-ABI: transcripts as UTF-32 text on the device, and the host-side packing of ragged logits.
//
// (1) coral_decoder_tokens_to_text: the winning token rows of coral_ctc_beam_decode -> one flat
//     code-point buffer + offsets, on the device. This is what pyctcdecode does with
//     "".join(alphabet[t] for t in ...) when it builds the beam text (UP:pyctcdecode decoder.py;
//     the text reaches the reference at R:src/coral/evaluate.py:61-73 and
//     R:src/coral/validation.py:121-134). Keeping the text on the device lets cer()/wer()
//     (R:src/coral/metrics.py:8-61 -> coral_edit_counts) consume the hypotheses without the
//     strings -> UTF-32 -> host-to-device round trip; the host receives the same flat buffer once
//     and only slices it into Python strings.
// (2) coral_host_pack_rows: the list of [T_i, V] arrays that
//     Wav2Vec2ProcessorWithLM.batch_decode hands over
//     (HF:models/wav2vec2_with_lm/processing_wav2vec2_with_lm.py:371, :398-406) is packed into ONE
//     pinned, ragged [sum T_i, V] buffer by a few host threads; the beam kernel then reads the
//     valid frames straight from it (frame_offsets_dev), so no padding is ever moved.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "common.cuh"
#include "handles.h"

namespace coral {

struct LabelTable {
  uint32_t cps[kVMax][kMaxLabelCps];
  uint8_t ncp[kVMax];
};

constexpr int kTextWarps = 8;

// one warp per utterance: number of code points of its transcript
__global__ void __launch_bounds__(kTextWarps * 32)
text_count_kernel(const __grid_constant__ LabelTable L, const uint8_t* __restrict__ tokens, int64_t row_pitch,
                  const int32_t* __restrict__ lens, int64_t lens_stride, int B, int64_t* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int u = blockIdx.x * kTextWarps + (threadIdx.x >> 5);
  if (u >= B) return;
  const int n = lens[(int64_t)u * lens_stride];
  const uint8_t* row = tokens + (int64_t)u * row_pitch;
  int c = 0;
  for (int i = lane; i < n; i += 32) c += L.ncp[row[i] & (kVMax - 1)];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) counts[u] = c;
}

// exclusive scan of counts[B] -> offsets[B + 1] (in place is fine: counts == offsets + 1 is NOT
// assumed), and the maximum count. One CTA: B is a batch of utterances, not a tensor.
__global__ void __launch_bounds__(1024)
text_scan_kernel(const int64_t* __restrict__ counts, int B, int64_t* __restrict__ offsets, int32_t* __restrict__ max_len) {
  __shared__ int64_t part[1024];
  __shared__ int32_t pmax[1024];
  const int t = threadIdx.x;
  const int per = (B + 1023) / 1024;
  const int a = min(B, t * per), b = min(B, a + per);
  int64_t s = 0;
  int32_t m = 0;
  for (int i = a; i < b; ++i) { s += counts[i]; m = max(m, (int32_t)counts[i]); }
  part[t] = s;
  pmax[t] = m;
  __syncthreads();
  // Hillis-Steele over 1024 partials
  for (int o = 1; o < 1024; o <<= 1) {
    const int64_t v = t >= o ? part[t - o] : 0;
    const int32_t w = t >= o ? pmax[t - o] : 0;
    __syncthreads();
    part[t] += v;
    pmax[t] = max(pmax[t], w);
    __syncthreads();
  }
  int64_t run = t ? part[t - 1] : 0;
  for (int i = a; i < b; ++i) { const int64_t c = counts[i]; offsets[i] = run; run += c; }
  if (t == 1023) { offsets[B] = part[1023]; if (max_len) *max_len = pmax[1023]; }
}

__global__ void __launch_bounds__(kTextWarps * 32)
text_write_kernel(const __grid_constant__ LabelTable L, const uint8_t* __restrict__ tokens, int64_t row_pitch,
                  const int32_t* __restrict__ lens, int64_t lens_stride, int B, const int64_t* __restrict__ offsets,
                  uint32_t* __restrict__ out, int64_t cap) {
  const int lane = threadIdx.x & 31;
  const int u = blockIdx.x * kTextWarps + (threadIdx.x >> 5);
  if (u >= B) return;
  const int n = lens[(int64_t)u * lens_stride];
  const uint8_t* row = tokens + (int64_t)u * row_pitch;
  int64_t base = offsets[u];
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    const int tok = i < n ? (row[i] & (kVMax - 1)) : 0;
    const int k = i < n ? L.ncp[tok] : 0;
    int incl = k;  // inclusive warp scan of the code-point counts
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const int64_t at = base + incl - k;
    for (int q = 0; q < k; ++q)
      if (at + q < cap) out[at + q] = L.cps[tok][q];
    base += __shfl_sync(0xffffffffu, incl, 31);
  }
}

}  // namespace coral

using namespace coral;

extern "C" {

int32_t coral_decoder_tokens_to_text(const coral_decoder* dec, const uint8_t* tokens_dev, int64_t row_pitch,
                                     const int32_t* lens_dev, int64_t lens_stride, int32_t B, uint32_t* out_cps_dev,
                                     int64_t cps_cap, int64_t* out_offsets_dev, int64_t* work_dev,
                                     int32_t* out_max_len_dev, void* stream) {
  if (!dec) return fail(CORAL_EARG, "coral_decoder_tokens_to_text: null decoder");
  if (B < 0 || row_pitch < 0 || cps_cap < 0) return fail(CORAL_EARG, "negative size");
  if (!out_offsets_dev) return fail(CORAL_EARG, "coral_decoder_tokens_to_text: null buffer");
  cudaStream_t st = (cudaStream_t)stream;
  DeviceGuard g(dec->device);
  if (B == 0) {
    CORAL_CUDA_OK(cudaMemsetAsync(out_offsets_dev, 0, sizeof(int64_t), st));
    if (out_max_len_dev) CORAL_CUDA_OK(cudaMemsetAsync(out_max_len_dev, 0, sizeof(int32_t), st));
    return CORAL_OK;
  }
  if (!tokens_dev || !lens_dev || !out_cps_dev || !work_dev) return fail(CORAL_EARG, "coral_decoder_tokens_to_text: null buffer");
  LabelTable L;
  memset(&L, 0, sizeof(L));
  for (int v = 0; v < dec->P.V; ++v) {
    L.ncp[v] = dec->P.label_ncp[v];
    for (int q = 0; q < kMaxLabelCps; ++q) L.cps[v][q] = dec->P.label_cps[v][q];
  }
  const unsigned grid = (unsigned)((B + kTextWarps - 1) / kTextWarps);
  text_count_kernel<<<grid, kTextWarps * 32, 0, st>>>(L, tokens_dev, row_pitch, lens_dev, lens_stride, B, work_dev);
  text_scan_kernel<<<1, 1024, 0, st>>>(work_dev, B, out_offsets_dev, out_max_len_dev);
  text_write_kernel<<<grid, kTextWarps * 32, 0, st>>>(L, tokens_dev, row_pitch, lens_dev, lens_stride, B,
                                                      out_offsets_dev, out_cps_dev, cps_cap);
  CORAL_CUDA_OK(cudaGetLastError());
  return CORAL_OK;
}

}  // extern "C"

namespace coral {
#if defined(__x86_64__) && !defined(__CUDA_ARCH__)
// Row copy with non-temporal stores: the destination is the pinned staging buffer, which nobody on
// the host reads again (the GPU pulls it over PCIe), so the stores need not allocate cache lines
// (no read-for-ownership: a third less memory traffic than a cached copy).
__attribute__((target("avx2"))) static void copy_row_stream(char* d, const char* s, size_t n) {
  const size_t head = (32 - (reinterpret_cast<uintptr_t>(d) & 31)) & 31;
  if (n < 256 + head) { memcpy(d, s, n); return; }
  memcpy(d, s, head);
  d += head; s += head; n -= head;
  size_t i = 0;
  for (; i + 128 <= n; i += 128) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 32));
    const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 64));
    const __m256i e = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 96));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i), a);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 32), b);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 64), c);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 96), e);
  }
  memcpy(d + i, s + i, n - i);
}
static const bool g_have_avx2 = __builtin_cpu_supports("avx2") && getenv("CORAL_B200_PACK_MEMCPY") == nullptr;
#endif
static inline void copy_row(char* d, const char* s, size_t n) {
#if defined(__x86_64__) && !defined(__CUDA_ARCH__)
  if (g_have_avx2) { copy_row_stream(d, s, n); return; }
#endif
  memcpy(d, s, n);
}
}  // namespace coral

extern "C" {

int32_t coral_host_pack_rows(const void* const* src, const int64_t* n_bytes, const int64_t* dst_offsets, int64_t n,
                             void* dst, int32_t n_threads) {
  if (n < 0) return fail(CORAL_EARG, "negative count");
  if (n == 0) return CORAL_OK;
  if (!src || !n_bytes || !dst_offsets || !dst) return fail(CORAL_EARG, "coral_host_pack_rows: null argument");
  int64_t total = 0;
  for (int64_t i = 0; i < n; ++i) total += n_bytes[i];
  // below ~1 MB per thread the thread start-up costs more than the copy
  int nt = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads > 0 ? n_threads : 1, total >> 20));
  auto work = [&](std::atomic<int64_t>* next) {
    for (;;) {
      const int64_t i0 = next->fetch_add(16);
      if (i0 >= n) break;
      const int64_t i1 = std::min<int64_t>(n, i0 + 16);
      for (int64_t i = i0; i < i1; ++i)
        if (n_bytes[i] > 0)
          copy_row(static_cast<char*>(dst) + dst_offsets[i], static_cast<const char*>(src[i]), (size_t)n_bytes[i]);
    }
#if defined(__x86_64__)
    _mm_sfence();  // the streaming stores are visible before the caller publishes the chunk
#endif
  };
  std::atomic<int64_t> next(0);
  if (nt == 1) {
    work(&next);
  } else {
    std::vector<std::thread> th;
    th.reserve(nt - 1);
    for (int k = 0; k < nt - 1; ++k) th.emplace_back(work, &next);
    work(&next);
    for (auto& t : th) t.join();
  }
  return CORAL_OK;
}

// The transcripts as a Python list of str, built in one C loop from the flat buffer that came
// back from the device (what pyctcdecode's decode_batch returns, UP:pyctcdecode decoder.py). The
// CPython entry points are looked up in the running process (no link-time dependency: the
// library still loads in a host without Python); the caller must hold the GIL, i.e. bind this
// symbol through ctypes.PyDLL. `kind` = bytes per symbol (1: Latin-1, 4: UTF-32). Returns a new
// reference, or NULL when no interpreter is present / an allocation failed.
void* coral_py_string_list(const void* data, int32_t kind, const int64_t* offsets, int64_t n) {
  typedef void* (*list_new_t)(ssize_t);
  typedef int (*list_set_t)(void*, ssize_t, void*);
  typedef void* (*from_kind_t)(int, const void*, ssize_t);
  typedef void (*decref_t)(void*);
  static list_new_t list_new = (list_new_t)dlsym(RTLD_DEFAULT, "PyList_New");
  static list_set_t list_set = (list_set_t)dlsym(RTLD_DEFAULT, "PyList_SetItem");
  static from_kind_t from_kind = (from_kind_t)dlsym(RTLD_DEFAULT, "PyUnicode_FromKindAndData");
  static decref_t decref = (decref_t)dlsym(RTLD_DEFAULT, "Py_DecRef");
  if (!list_new || !list_set || !from_kind || !decref || !offsets || n < 0 || (kind != 1 && kind != 4)) return nullptr;
  void* list = list_new((ssize_t)n);
  if (!list) return nullptr;
  const char* base = static_cast<const char*>(data);
  for (int64_t i = 0; i < n; ++i) {
    void* s = from_kind(kind, base + offsets[i] * kind, (ssize_t)(offsets[i + 1] - offsets[i]));
    if (!s) { decref(list); return nullptr; }
    list_set(list, (ssize_t)i, s);  // steals the reference
  }
  return list;
}

// Py_buffer as CPython 3.x lays it out (Include/pybuffer.h / object.h; unchanged since 3.0)
struct CoralPyBuffer {
  void* buf;
  void* obj;
  ssize_t len;
  ssize_t itemsize;
  int readonly;
  int ndim;
  char* format;
  ssize_t* shape;
  ssize_t* strides;
  ssize_t* suboffsets;
  void* internal;
};

int64_t coral_py_logits_rows(void* list, int32_t V, int64_t* out_ptrs, int64_t* out_frames, uint8_t* out_other) {
  typedef ssize_t (*list_size_t)(void*);
  typedef void* (*list_get_t)(void*, ssize_t);
  typedef int (*get_buffer_t)(void*, CoralPyBuffer*, int);
  typedef void (*release_t)(CoralPyBuffer*);
  typedef void (*err_clear_t)(void);
  typedef int (*list_check_t)(void*);
  static list_size_t list_size = (list_size_t)dlsym(RTLD_DEFAULT, "PyList_Size");
  static list_get_t list_get = (list_get_t)dlsym(RTLD_DEFAULT, "PyList_GetItem");
  static get_buffer_t get_buffer = (get_buffer_t)dlsym(RTLD_DEFAULT, "PyObject_GetBuffer");
  static release_t release = (release_t)dlsym(RTLD_DEFAULT, "PyBuffer_Release");
  static err_clear_t err_clear = (err_clear_t)dlsym(RTLD_DEFAULT, "PyErr_Clear");
  if (!list_size || !list_get || !get_buffer || !release || !err_clear || !list || !out_ptrs || !out_frames ||
      !out_other || V <= 0)
    return -1;
  const ssize_t n = list_size(list);
  if (n < 0) { err_clear(); return -1; }
  constexpr int kStridesFormat = 0x0008 | 0x0010 | 0x0004;  // PyBUF_STRIDES (incl. PyBUF_ND) | PyBUF_FORMAT
  for (ssize_t i = 0; i < n; ++i) {
    void* item = list_get(list, i);  // borrowed
    CoralPyBuffer view;
    bool ok = false;
    if (item && get_buffer(item, &view, kStridesFormat) == 0) {
      const char* f = view.format ? view.format : "";
      if (*f == '<' || *f == '=' || *f == '@') ++f;
      ok = view.ndim == 2 && view.itemsize == 4 && f[0] == 'f' && f[1] == 0 && view.shape && view.strides &&
           view.shape[1] == V && (view.shape[0] == 0 || (view.strides[1] == 4 && (view.shape[0] == 1 || view.strides[0] == 4 * (ssize_t)V)));
      if (ok) {
        out_ptrs[i] = (int64_t)(intptr_t)view.buf;
        out_frames[i] = (int64_t)view.shape[0];
      }
      release(&view);
    } else {
      err_clear();
    }
    out_other[i] = ok ? 0 : 1;
    if (!ok) { out_ptrs[i] = 0; out_frames[i] = 0; }
  }
  return (int64_t)n;
}

}  // extern "C"
