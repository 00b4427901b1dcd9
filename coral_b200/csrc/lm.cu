// C-ABI: n-gram LM handle -- ARPA load, upload to HBM, debug scoring kernel.
// Replaces kenlm.Model for the decode path (see include/coral_b200.h, lm_tables.h).
#include <vector>

#include "common.cuh"
#include "handles.h"

namespace coral {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int32_t fail(int32_t code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
int sm_count(int device) {
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) n = 148;
  return n;
}

// One thread per sentence: runs the KenLM state machine over its words on the device
// tables. A parity/debug kernel (A8), not a hot path.
__global__ void lm_score_sentences_kernel(LmView lm, const uint32_t* __restrict__ cps,
                                          const int64_t* __restrict__ word_off,
                                          const int64_t* __restrict__ sent_off, int64_t n_sent, int bos, int eos,
                                          float* __restrict__ out_prob, int32_t* __restrict__ out_oov) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_sent) return;
  LmState st, nx;
  if (bos) lm_begin_sentence(lm, st); else lm_null_context(st);
  for (int64_t w = sent_off[s]; w < sent_off[s + 1]; ++w) {
    uint64_t h = kWordHashSeed;
    for (int64_t q = word_off[w]; q < word_off[w + 1]; ++q) h = word_hash_push(h, cps[q]);
    uint32_t wid = 0, fl = 0;
    const bool in_lm = lex_find(lm, h, wid, fl) && (fl & kLexInLm);
    if (!in_lm) wid = 0;
    out_prob[w + (eos ? s : 0)] = lm_base_score(lm, st, wid, nx);
    out_oov[w] = in_lm ? 0 : 1;
    st = nx;
  }
  if (eos) out_prob[sent_off[s + 1] + s] = lm_base_score(lm, st, lm.eos_id, nx);
}

}  // namespace coral

using namespace coral;

extern "C" {

const char* coral_last_error(void) { return g_last_error.c_str(); }
int32_t coral_abi_version(void) { return 1; }

static int32_t lm_load(const char* path, int32_t device, coral_lm** out, int kind);

int32_t coral_lm_load_arpa(const char* path, int32_t device, coral_lm** out) { return lm_load(path, device, out, 0); }
int32_t coral_lm_load_kenlm_binary(const char* path, int32_t device, coral_lm** out) { return lm_load(path, device, out, 1); }
int32_t coral_lm_load(const char* path, int32_t device, coral_lm** out) {
  if (!path || !out) return fail(CORAL_EARG, "coral_lm_load: null argument");
  return lm_load(path, device, out, is_kenlm_binary(path) ? 1 : 0);
}

static int32_t lm_load(const char* path, int32_t device, coral_lm** out, int kind) {
  if (!path || !out) return fail(CORAL_EARG, "coral_lm_load: null argument");
  *out = nullptr;
  coral_lm* lm = new coral_lm();
  std::string err;
  int rc = kind == 1 ? load_kenlm_binary(path, lm->host, err) : load_arpa(path, lm->host, err);
  if (rc != 0) { delete lm; return fail(rc, err); }
  rc = build_lexicon(lm->host, nullptr, lm->vocab_lex, err);
  if (rc != 0) { delete lm; return fail(rc, err); }
  lm->device = device;
  DeviceGuard g(device);
  if (!g.ok) { delete lm; return fail(CORAL_ECUDA, "cannot select CUDA device"); }
  const size_t ub = lm->host.uni.size() * sizeof(UniEntry);
  const size_t nb = lm->host.ng.size() * sizeof(NgSlot);
  const size_t lb = lm->vocab_lex.lex.size() * sizeof(LexSlot);
  cudaError_t e;
  if ((e = cudaMalloc(&lm->d_uni, ub)) != cudaSuccess || (e = cudaMalloc(&lm->d_ng, nb)) != cudaSuccess ||
      (e = cudaMalloc(&lm->d_lex, lb)) != cudaSuccess ||
      (e = cudaMemcpy(lm->d_uni, lm->host.uni.data(), ub, cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(lm->d_ng, lm->host.ng.data(), nb, cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(lm->d_lex, lm->vocab_lex.lex.data(), lb, cudaMemcpyHostToDevice)) != cudaSuccess) {
    std::string m = std::string("uploading LM tables: ") + cudaGetErrorString(e);
    coral_lm_free(lm);
    return fail(CORAL_ECUDA, m);
  }
  lm->device_bytes = ub + nb + lb;
  *out = lm;
  return CORAL_OK;
}

int32_t coral_lm_free(coral_lm* lm) {
  if (!lm) return CORAL_OK;
  DeviceGuard g(lm->device);
  if (lm->d_uni) cudaFree(lm->d_uni);
  if (lm->d_ng) cudaFree(lm->d_ng);
  if (lm->d_lex) cudaFree(lm->d_lex);
  delete lm;
  return CORAL_OK;
}

int32_t coral_lm_info(const coral_lm* lm, int32_t* order, uint64_t* ngram_counts, uint64_t* vocab_size,
                      uint64_t* device_bytes) {
  if (!lm) return fail(CORAL_EARG, "coral_lm_info: null handle");
  if (order) *order = lm->host.order;
  if (ngram_counts)
    for (int i = 0; i < lm->host.order; ++i) ngram_counts[i] = lm->host.loaded[i];
  if (vocab_size) *vocab_size = lm->host.uni.size();
  if (device_bytes) *device_bytes = lm->device_bytes;
  return CORAL_OK;
}

int32_t coral_lm_contains(const coral_lm* lm, const uint32_t* word_cps, const int64_t* word_offsets,
                          int64_t n_words, int32_t* out) {
  if (!lm || !word_offsets || !out) return fail(CORAL_EARG, "coral_lm_contains: null argument");
  LmView v = make_view(lm->host, lm->vocab_lex, lm->host.uni.data(), lm->host.ng.data(), lm->vocab_lex.lex.data());
  for (int64_t i = 0; i < n_words; ++i) {
    uint64_t h = kWordHashSeed;
    for (int64_t q = word_offsets[i]; q < word_offsets[i + 1]; ++q) h = word_hash_push(h, word_cps[q]);
    uint32_t wid = 0, fl = 0;
    out[i] = (word_offsets[i + 1] > word_offsets[i] && lex_find(v, h, wid, fl) && (fl & kLexInLm)) ? 1 : 0;
  }
  return CORAL_OK;
}

int32_t coral_lm_score_sentences(const coral_lm* lm, const uint32_t* word_cps_dev, const int64_t* word_offsets_dev,
                                 const int64_t* sent_offsets_dev, int64_t n_sentences, int32_t bos, int32_t eos,
                                 float* out_probs_dev, int32_t* out_oov_dev, void* stream) {
  if (!lm || !word_offsets_dev || !sent_offsets_dev || !out_probs_dev || !out_oov_dev)
    return fail(CORAL_EARG, "coral_lm_score_sentences: null argument");
  if (n_sentences <= 0) return CORAL_OK;
  DeviceGuard g(lm->device);
  LmView v = make_view(lm->host, lm->vocab_lex, lm->d_uni, lm->d_ng, lm->d_lex);
  const int threads = 128;
  const int64_t blocks = (n_sentences + threads - 1) / threads;
  lm_score_sentences_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
      v, word_cps_dev, word_offsets_dev, sent_offsets_dev, n_sentences, bos, eos, out_probs_dev, out_oov_dev);
  CORAL_CUDA_OK(cudaGetLastError());
  return CORAL_OK;
}

}  // extern "C"
