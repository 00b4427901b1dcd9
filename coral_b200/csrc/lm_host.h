// Host-side loader: ARPA text -> the open-addressing tables of lm_tables.h.
// Plain C++ (no CUDA) so that the library and the tests' host simulation share it.
//
// On-disk format accepted: the ARPA files lmplz writes, including the shape CoRal's
// patch produces (R:src/coral/ngram.py:147-169: "ngram 1=" bumped by one and a second
// "</s>" unigram line copied from "<s>"); a duplicate unigram keeps its first slot.
#pragma once
#include <stdint.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "lm_tables.h"

namespace coral {

struct HostLm {
  std::string path;
  int order = 0;
  std::vector<uint64_t> counts;  // declared "ngram n=" values
  std::vector<uint64_t> loaded;  // n-grams actually stored per order
  std::vector<std::string> words;  // id -> utf-8
  std::unordered_map<std::string, uint32_t> vocab;
  std::vector<UniEntry> uni;
  std::vector<NgSlot> ng;
  uint64_t ng_mask = 0;
  uint32_t bos_id = 0, eos_id = 0;
  int kenlm_keys = 0;  // ng holds KenLM chain keys (read from a KenLM binary), see lm_tables.h
  float score_ub = 0.0f;  // no BaseScore result exceeds this log10 value (>= 0; see compute_score_ub)
};

struct HostLexicon {
  std::vector<LexSlot> lex;
  uint64_t lex_mask = 0;
  int has_unigrams = 0;
  uint64_t n_entries = 0;
  // per slot of `lex`: alphabet tokens that extend the prefix penalty-free (see LmView::lex_ok);
  // empty when build_lexicon was not given the alphabet
  std::vector<uint64_t> child_ok;
  uint64_t root_ok = 0;
};

// returns 0 on success; negative status + message otherwise
int load_arpa(const char* path, HostLm& lm, std::string& err);

// KenLM probing binary (what CoRal ships as language_model/{N}gram.bin, R:src/coral/ngram.py:361-387)
int load_kenlm_binary(const char* path, HostLm& lm, std::string& err);
// true if the file starts with KenLM's binary magic
bool is_kenlm_binary(const char* path);

// unigrams == nullptr <=> pyctcdecode's ``unigrams=None`` (no unigram set, no char trie)
// labels != nullptr: also fill HostLexicon::child_ok / root_ok for that alphabet (<= 64 labels)
int build_lexicon(const HostLm& lm, const std::vector<std::u32string>* unigrams, HostLexicon& out,
                  std::string& err, const std::vector<std::u32string>* labels = nullptr);

bool utf8_to_u32(const std::string& s, std::u32string& out);
uint64_t hash_word(const std::u32string& w);

inline LmView make_view(const HostLm& lm, const HostLexicon& lx, const UniEntry* uni,
                        const NgSlot* ng, const LexSlot* lex, const uint64_t* lex_ok = nullptr) {
  LmView v;
  v.uni = uni;
  v.ng = ng;
  v.lex = lex;
  v.ng_mask = lm.ng_mask;
  v.lex_mask = lx.lex_mask;
  v.n_vocab = (uint32_t)lm.uni.size();
  v.order = lm.order;
  v.bos_id = lm.bos_id;
  v.eos_id = lm.eos_id;
  v.has_unigrams = lx.has_unigrams;
  v.kenlm_keys = lm.kenlm_keys;
  v.score_ub = lm.score_ub;
  v.lex_ok = lx.child_ok.empty() ? nullptr : lex_ok;
  v.root_ok = lx.root_ok;
  v.present = 1;
  return v;
}

}  // namespace coral
