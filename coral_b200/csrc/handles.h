// Opaque handle definitions shared by the C-ABI translation units.
#pragma once
#include <map>
#include <mutex>
#include <stdint.h>

#include "beam_core.h"
#include "lm_host.h"

struct coral_lm {
  coral::HostLm host;
  coral::HostLexicon vocab_lex;  // LM vocabulary only (used by the LM debug entry points)
  coral::UniEntry* d_uni = nullptr;
  coral::NgSlot* d_ng = nullptr;
  coral::LexSlot* d_lex = nullptr;
  int device = 0;
  uint64_t device_bytes = 0;
};

struct coral_decoder {
  const coral_lm* lm = nullptr;
  coral::HostLexicon lex;  // LM vocabulary U unigram list, with pyctcdecode's flags
  coral::LexSlot* d_lex = nullptr;
  uint64_t* d_lex_ok = nullptr;  // per lexicon slot: tokens that extend the prefix penalty-free
  coral::DecodeParams P;   // alphabet + current alpha/beta/unk/boundary
  int device = 0;
  // scratch arenas in HBM, one set per CUDA stream the decoder is used on (launches on
  // different streams may overlap), grown lazily, reused call after call
  struct Scratch {
    uint8_t* d_scratch = nullptr;
    size_t scratch_bytes = 0;
    size_t slot_bytes = 0;
    uint32_t n_slots = 0;
    uint32_t node_cap = 0, bnd_cap = 0, ch_size = 0, outs_cap = 0, wf_cap = 0, hist_cap = 0;
    int32_t* d_work = nullptr;   // 4 counters + the heavy kernel's queue
    size_t work_cap = 0;         // ints allocated at d_work
    bool budget_capped = false;  // the memory budget, not the launch, limited n_slots
  };
  std::map<void*, Scratch> scratch;
  std::mutex mu;
  uint64_t device_bytes = 0;
};
