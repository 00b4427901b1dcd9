/* coral_b200 -- C ABI of the B200-native CTC-decode + WER/CER hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types. Every entry
 * point cites the reference interface it replaces (R: = alexandrainst/coral,
 * HF: = transformers 5.5.0, UP: = un-vendored upstream named in SURVEY.md section 8).
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns an int32 status: 0 OK, -1 bad argument/shape, -2 I/O or
 *     parse error, -3 CUDA error, -4 capacity exceeded. coral_last_error() returns the
 *     message for the calling thread. No exception crosses this boundary.
 *   - pointers named *_dev are device pointers owned by the caller (e.g. torch tensors);
 *     they are borrowed until the work queued on `stream` completes. `stream` is a
 *     cudaStream_t passed as void* (NULL = default stream). All kernels are queued
 *     asynchronously; the caller synchronises.
 *   - handles (coral_lm, coral_decoder) are owned by the library, immutable after
 *     creation except coral_decoder_set_params (caller serialises, as HF calls
 *     reset_params before every decode). Per-decoder scratch in HBM grows lazily.
 *   - strings cross the boundary as UTF-32 code points + int64 offsets ([n+1]).
 */
#ifndef CORAL_B200_H_
#define CORAL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct coral_lm coral_lm;
typedef struct coral_decoder coral_decoder;
typedef struct coral_normaliser coral_normaliser;

#define CORAL_OK 0
#define CORAL_EARG (-1)
#define CORAL_EIO (-2)
#define CORAL_ECUDA (-3)
#define CORAL_ECAP (-4)

#define CORAL_MAX_ORDER 6
#define CORAL_MAX_VOCAB 64 /* alphabet entries (CoRal: 46) */

const char* coral_last_error(void);
int32_t coral_abi_version(void);

/* ---------------------------------------------------------------- n-gram LM (A6/A8) */

/* kenlm.Model(path) for an ARPA file (UP:kenlm python/kenlm.pyx; reached from
 * R:src/coral/ngram.py:341-343 via pyctcdecode.build_ctcdecoder(kenlm_model_path=...)).
 * Parses on the host and uploads the tables to `device`. */
int32_t coral_lm_load_arpa(const char* path, int32_t device, coral_lm** out);
/* The same from a KenLM *probing* binary (what CoRal ships as language_model/{N}gram.bin,
 * R:src/coral/ngram.py:361-387). The reader is a restatement of the published format that no real
 * build_binary output was available to check: it accepts a file only if every
 * structural invariant holds and refuses trie / quantised models by name (CORAL_EIO). */
int32_t coral_lm_load_kenlm_binary(const char* path, int32_t device, coral_lm** out);
/* kenlm.Model(path): ARPA or KenLM binary, told apart by the magic bytes. */
int32_t coral_lm_load(const char* path, int32_t device, coral_lm** out);
int32_t coral_lm_free(coral_lm* lm);
/* kenlm.Model.order, n-gram counts per order [order], vocabulary size, HBM bytes. */
int32_t coral_lm_info(const coral_lm* lm, int32_t* order, uint64_t* ngram_counts, uint64_t* vocab_size,
                      uint64_t* device_bytes);
/* `word in kenlm_model` (UP:kenlm.pyx Model.__contains__): out[i] = vocabulary index != 0. */
int32_t coral_lm_contains(const coral_lm* lm, const uint32_t* word_cps, const int64_t* word_offsets,
                          int64_t n_words, int32_t* out);
/* kenlm.Model.full_scores-style check of BaseScore on the DEVICE tables: sentences of
 * words; out_probs gets one float32 log10 prob per word (+1 per sentence when eos),
 * laid out at word_index + sentence_index*eos. Used by the parity tests of A8. */
int32_t coral_lm_score_sentences(const coral_lm* lm, const uint32_t* word_cps_dev,
                                 const int64_t* word_offsets_dev, const int64_t* sent_offsets_dev,
                                 int64_t n_sentences, int32_t bos, int32_t eos, float* out_probs_dev,
                                 int32_t* out_oov_dev, void* stream);

/* ------------------------------------------------------------------- decoder (A5-A7) */

/* pyctcdecode.build_ctcdecoder(labels, kenlm_model_path, unigrams, ...) after
 * Alphabet.build_alphabet (R:src/coral/ngram.py:341-343). `labels` are the
 * NORMALISED alphabet entries ("" blank, " " word delimiter) as UTF-32; lm may be NULL
 * (no-LM mode); n_unigrams < 0 means unigrams=None. */
int32_t coral_decoder_create(const uint32_t* label_cps, const int32_t* label_offsets, int32_t n_labels,
                             int32_t blank_id, int32_t space_id, const coral_lm* lm,
                             const uint32_t* unigram_cps, const int64_t* unigram_offsets, int64_t n_unigrams,
                             int32_t device, coral_decoder** out);
int32_t coral_decoder_free(coral_decoder* dec);
/* BeamSearchDecoderCTC.reset_params / LanguageModel attributes
 * (HF:models/wav2vec2_with_lm/processing_wav2vec2_with_lm.py:365-367, :160-183). */
int32_t coral_decoder_set_params(coral_decoder* dec, double alpha, double beta, double unk_score_offset,
                                 int32_t score_boundary);
/* HBM bytes held by the decoder (lexicon + scratch). */
int32_t coral_decoder_info(const coral_decoder* dec, uint64_t* lexicon_entries, uint64_t* device_bytes);

/* BeamSearchDecoderCTC.decode_beams_batch / decode_batch over a padded batch
 * (UP:pyctcdecode decoder.py; HF:...processing_wav2vec2_with_lm.py:398-406, :565-572;
 * HF:pipelines/automatic_speech_recognition.py:612-616).
 *   logits_dev   float32 [B, T_max, V]; rows >= lengths[b] are ignored
 *   lengths_dev  int32 [B]
 *   order_dev    int32 [B] processing order (longest first balances the tail) or NULL
 *   frame_offsets_dev  int64 [B] or NULL. Ragged input: logits_dev is a packed [sum T, V] buffer
 *                and utterance b starts at frame frame_offsets_dev[b] (what the list of [T_i, V]
 *                arrays of HF:...processing_wav2vec2_with_lm.py:371 becomes without moving any
 *                padding). T_max is then only the pitch of out_tokens (>= max lengths).
 *                logits_dev may also be PINNED HOST memory (cudaHostAlloc / torch pin_memory:
 *                device-accessible under unified addressing): the kernel then pulls each valid
 *                frame over PCIe exactly once, with no staging copy -- the end-to-end path.
 *   input_mode   0 = pyctcdecode's auto-detection of probabilities vs logits (per utterance,
 *                inside the kernel), 1 logits, 2 probs
 *   n_best       beams returned per utterance (<= beam_width)
 * Outputs (device, caller-allocated):
 *   out_n_beams     int32 [B]           number of final beams (<= beam_width)
 *   out_logit_score float64 [B, n_best]
 *   out_lm_score    float64 [B, n_best] combined score (pyctcdecode's "lm_score")
 *   out_tokens      uint8 [B, n_best, T_max] alphabet indices of the text (no blanks)
 *   out_lens        int32 [B, n_best]
 *   out_status      int32 [B]           0, CORAL_ECAP (arena capacity) or CORAL_ECUDA (streamed
 *                                       input never arrived) for that utterance
 *   stats_dev       uint64 [32] or NULL: beam extensions, LM word scorings, n-gram probes,
 *                   frames, lexicon probes, back-pointer records, LM boundary records, frames that
 *                   needed the radix select (work counters of SURVEY 8d); [8..15] = cycles per
 *                   kernel phase, [16..23] / [24..31] = cycles / calls of selected device
 *                   operations (tuning). A non-NULL buffer selects the instrumented
 *                   instantiation of the kernel, which is about twice as slow: pass NULL unless
 *                   the counters are wanted.
 *   ready_dev       int32 device scalar or NULL. Streamed input: the kernel may be launched
 *                   while the logits are still being copied in utterance order on ANOTHER
 *                   stream, ready_chunk utterances at a time; after each chunk the copier stores
 *                   the number of utterances delivered so far into *ready_dev (a 4-byte copy on
 *                   the same copy stream). A thread group waits for its utterance's chunk. With
 *                   a ready counter, order_dev must list chunk 0's utterances first, then chunk
 *                   1's, ... (any order inside a chunk).
 *   out_word_frames_dev  int32 [B, n_best, max_words, 2] or NULL: pyctcdecode's text_frames, the
 *                   (start, end) frame of every word of each returned beam (what HF turns into
 *                   word_offsets, HF:...processing_wav2vec2_with_lm.py:416-443); with it,
 *   out_word_counts_dev  int32 [B, n_best] words written per beam. max_words >= (T_max + 1) / 2 + 1
 *                   holds every possible transcript. NULL selects the kernel without word timing.
 * prune_history != 0 is pyctcdecode's prune_history=True: after every frame's trim only the best
 * beam per (last max(1, LM order - 1) words of the text, word_part, last_char) stays.
 * Hotwords are not implemented: pyctcdecode's HotwordScorer scores partial words through
 * next(CharTrie.iterkeys(prefix, shallow=True)) over a trie built from a Python set, i.e. its
 * result depends on set iteration order (not reproducible across processes), and CoRal never
 * passes hotwords; the Python surface raises NotImplementedError for them. */
int32_t coral_ctc_beam_decode(coral_decoder* dec, const float* logits_dev, const int32_t* lengths_dev,
                              const int32_t* order_dev, const int64_t* frame_offsets_dev, int32_t B,
                              int32_t T_max, int32_t V, int32_t beam_width, double beam_prune_logp, double token_min_logp,
                              int32_t prune_history, int32_t input_mode, int32_t n_best,
                              int32_t* out_n_beams_dev, double* out_logit_score_dev, double* out_lm_score_dev,
                              uint8_t* out_tokens_dev, int32_t* out_lens_dev, int32_t* out_status_dev,
                              uint64_t* stats_dev, const int32_t* ready_dev, int32_t ready_chunk,
                              int32_t* out_word_frames_dev, int32_t* out_word_counts_dev, int32_t max_words,
                              void* stream);

/* The text of returned beams as UTF-32 ON THE DEVICE: what pyctcdecode builds with
 * "".join(labels[t] ...) for the beam it returns (UP:pyctcdecode decoder.py), kept where the
 * metric kernels can read it (R:src/coral/metrics.py:8-61 -> coral_edit_counts) so that the
 * hypotheses never make the strings -> UTF-32 -> host-to-device round trip.
 *   tokens_dev   uint8 rows of alphabet indices; row b starts at tokens_dev + b * row_pitch
 *                (beam 0 of out_tokens [B, n_best, T_max]: row_pitch = n_best * T_max)
 *   lens_dev     int32, length of row b at lens_dev[b * lens_stride]
 *   out_cps_dev  uint32 [cps_cap] flat code points (cps_cap >= sum of lengths x longest label)
 *   out_offsets_dev int64 [B + 1]
 *   work_dev     int64 [B] scratch
 *   out_max_len_dev int32 scalar or NULL: longest transcript in code points */
int32_t coral_decoder_tokens_to_text(const coral_decoder* dec, const uint8_t* tokens_dev, int64_t row_pitch,
                                     const int32_t* lens_dev, int64_t lens_stride, int32_t B,
                                     uint32_t* out_cps_dev, int64_t cps_cap, int64_t* out_offsets_dev,
                                     int64_t* work_dev, int32_t* out_max_len_dev, void* stream);

/* HOST helper: copies n byte ranges src[i][0 .. n_bytes[i]) to dst + dst_offsets[i] with
 * n_threads host threads. Packs the list of [T_i, V] float32 arrays that
 * Wav2Vec2ProcessorWithLM.batch_decode hands to decode_beams_batch
 * (HF:models/wav2vec2_with_lm/processing_wav2vec2_with_lm.py:371, :398-406) into one pinned ragged
 * buffer for coral_ctc_beam_decode(frame_offsets_dev). Pure host code, no CUDA call. */
int32_t coral_host_pack_rows(const void* const* src, const int64_t* n_bytes, const int64_t* dst_offsets,
                             int64_t n, void* dst, int32_t n_threads);

/* HOST helper for Python hosts: n strings out of one flat buffer (kind = bytes per symbol: 1 Latin-1,
 * 4 UTF-32; offsets [n + 1] in symbols) as a new Python list of str -- the list decode_batch returns
 * (UP:pyctcdecode decoder.py). Resolves the CPython API in the running process at call time, so
 * the library has no link-time dependency on Python; call it with the GIL held (ctypes.PyDLL).
 * Returns a PyObject* (new reference) or NULL. */
void* coral_py_string_list(const void* data, int32_t kind, const int64_t* offsets, int64_t n);

/* HOST helper for Python hosts: the row pointers of a Python list of [T_i, V] float32 C-contiguous
 * buffers (numpy arrays) -- what Wav2Vec2ProcessorWithLM.batch_decode passes to decode_beams_batch
 * (HF:models/wav2vec2_with_lm/processing_wav2vec2_with_lm.py:371, :398-406) -- read through the buffer
 * protocol in one C loop instead of a Python loop over 8192 arrays. out_ptrs / out_frames [n] get the
 * data address and T_i of every conforming item; out_other[i] = 1 marks an item the caller must convert
 * itself (not a buffer, wrong dtype / rank / width, not contiguous). The list keeps the arrays alive;
 * nothing is retained. Returns n, or -1. Call with the GIL held (ctypes.PyDLL). */
int64_t coral_py_logits_rows(void* list, int32_t V, int64_t* out_ptrs, int64_t* out_frames, uint8_t* out_other);

/* -------------------------------------------------------------------- greedy (A3/A4) */

/* np.argmax(axis=-1) + Wav2Vec2CTCTokenizer grouping (R:src/coral/compute_metrics.py:62-70;
 * HF:models/wav2vec2/tokenization_wav2vec2.py:296-357, :410-459).
 *   logits_dev     float32 [B, T_max, V]
 *   lengths_dev    int32 [B] or NULL (= T_max for all)
 *   pad_fixup      1: a frame whose V logits all equal -100 decodes to blank_id
 *                  (R:src/coral/compute_metrics.py:66)
 *   out_ids_dev    int32 [B, T_max] argmax per frame (first maximum wins) or NULL
 *   out_tokens_dev int32 [B, T_max] ids after collapsing repeats and dropping blank_id
 *   out_lens_dev   int32 [B] */
int32_t coral_ctc_greedy(const float* logits_dev, const int32_t* lengths_dev, int32_t B, int32_t T_max,
                         int32_t V, int32_t blank_id, int32_t pad_fixup, int32_t* out_ids_dev,
                         int32_t* out_tokens_dev, int32_t* out_lens_dev, void* stream);
/* Collapse already-decoded ids (the 2-D branch / label decoding with group_tokens on/off). */
int32_t coral_ctc_collapse(const int32_t* ids_dev, const int32_t* lengths_dev, int32_t B, int32_t T_max,
                           int32_t blank_id, int32_t group_tokens, int32_t* out_tokens_dev,
                           int32_t* out_lens_dev, void* stream);

/* ---------------------------------------------------------- edit counts (A1/A2/A11/A12) */

#define CORAL_EDIT_TOKENS 0 /* sequences compared as given */
#define CORAL_EDIT_CHARS 1  /* jiwer cer_default: Strip -> list of characters */
#define CORAL_EDIT_WORDS 2  /* jiwer wer_default: RemoveMultipleSpaces -> Strip -> split(" ") */

/* jiwer.process_characters / process_words -> rapidfuzz Levenshtein.editops counts
 * (R:src/coral/metrics.py:26-33, :54-61).
 *   out_sdih_dev   int32 [n_pairs, 4] = substitutions, deletions, insertions, hits
 *   out_status_dev int32 [n_pairs]   0; 1 = the reference is empty after the transform (counts are then
 *                  S = D = H = 0, I = len(hyp): what jiwer >= 3.1 returns; jiwer 3.0 raised ValueError --
 *                  the caller chooses, see coral_b200/metrics.py); 2 = a string exceeds max_len, or the
 *                  pair is outside rapidfuzz's direct-alignment range (len1 * len2 >= 2^22 symbols:
 *                  rapidfuzz switches to Hirschberg splitting there) -- its counts are zeroed, never
 *                  computed on truncated input
 *   max_len        upper bound on the code points of any one string (sizes the work areas;
 *                  <= 2048, larger values return CORAL_ECAP)
 * Launches on different streams use separate off-chip work areas (one per device and stream). */
int32_t coral_edit_counts(const uint32_t* ref_cps_dev, const int64_t* ref_offsets_dev,
                          const uint32_t* hyp_cps_dev, const int64_t* hyp_offsets_dev, int64_t n_pairs,
                          int32_t mode, int64_t max_len, int32_t device, int32_t* out_sdih_dev,
                          int32_t* out_status_dev, void* stream);

/* Same, with explicit [begin, end) spans per string (e.g. hypotheses still sitting in the
 * padded decoder output on the device: begin = row * pitch, end = begin + length). */
int32_t coral_edit_counts_spans(const uint32_t* ref_cps_dev, const int64_t* ref_begin_dev,
                                const int64_t* ref_end_dev, const uint32_t* hyp_cps_dev,
                                const int64_t* hyp_begin_dev, const int64_t* hyp_end_dev, int64_t n_pairs,
                                int32_t mode, int64_t max_len, int32_t device, int32_t* out_sdih_dev,
                                int32_t* out_status_dev, void* stream);

/* --------------------------------------------------------- text normaliser */

/* HOST code. The text part of coral.data.process_example (R:src/coral/data.py:658-701) as called
 * between decoding and scoring with audio_column=None (R:src/coral/evaluate.py:61-72,
 * R:src/coral/validation.py:121-132), incl. coral.utils.convert_numeral_to_words
 * (R:src/coral/utils.py:303-472): numerals -> words, lower, filler words, NFKC, conversion dict (in
 * order), characters_to_keep (case-insensitive), space collapsing, line stripping.
 *   keep_cps / n_keep      characters_to_keep as code points; n_keep < 0 means None (keep all)
 *   conv_cps, conv_offsets conversion_dict.items() in order: key0, value0, key1, value1, ... as one
 *                          UTF-32 buffer with 2 * n_conv + 1 offsets */
int32_t coral_normaliser_create(const uint32_t* keep_cps, int64_t n_keep, const uint32_t* conv_cps,
                                const int64_t* conv_offsets, int64_t n_conv, int32_t lower_case,
                                int32_t convert_numerals, coral_normaliser** out);
int32_t coral_normaliser_free(coral_normaliser* h);
/* Normalises n strings (UTF-32 + offsets [n + 1]) with n_threads host threads; the results stay in
 * the handle. out_total = code points of the results. */
int32_t coral_normaliser_run(coral_normaliser* h, const uint32_t* cps, const int64_t* offsets, int64_t n,
                             int32_t n_threads, int64_t* out_total);
/* Copies the results of the last run: out_cps [total], out_offsets [n + 1], out_status [n]
 * (0 = normalised; 1 = holds something this restatement refuses instead of approximating: a Greek
 * capital sigma under lower_case, a non-ASCII decimal digit under convert_numerals). */
int32_t coral_normaliser_fetch(const coral_normaliser* h, uint32_t* out_cps, int64_t* out_offsets,
                               int32_t* out_status);
/* Unicode version of the tables compiled in (from the interpreter that ran the build). */
const char* coral_normaliser_unicode_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CORAL_B200_H_ */
