"""pyctcdecode ``BeamSearchDecoderCTC`` restated (test infrastructure only).

Behavioural restatement of UP:pyctcdecode 0.5.0 ``decoder.py`` and ``alphabet.py``
 with the same data structures as upstream -- beam
tuples, ``dict`` merge, text-keyed LM cache, ``heapq.nlargest`` -- so that its
ordering semantics are upstream's by construction. The reference reaches this code
at R:src/coral/ngram.py:341-343 (``build_ctcdecoder``), through
HF:models/wav2vec2_with_lm/processing_wav2vec2_with_lm.py:398-406
(``decode_beams_batch``) and :565-572 (``decode_beams``), and through
HF:pipelines/automatic_speech_recognition.py:612-616. Parity unpinned -- see
``oracle/__init__.py``.

Documented deviation (SURVEY A5 note ii): upstream iterates ``idx_list`` as a
CPython ``set`` of ``np.int64``; here tokens are visited in ascending id. The
order can only matter when two distinct beams tie bit-for-bit in float64.
"""

from __future__ import annotations

import heapq
import math
import re

import numpy as np

from .arpa import ArpaModel, load_unigram_set_from_arpa
from .lm import (
    DEFAULT_ALPHA,
    DEFAULT_BEAM_WIDTH,
    DEFAULT_BETA,
    DEFAULT_HOTWORD_WEIGHT,
    DEFAULT_MIN_TOKEN_LOGP,
    DEFAULT_PRUNE_LOGP,
    DEFAULT_SCORE_LM_BOUNDARY,
    DEFAULT_UNK_LOGP_OFFSET,
    MIN_TOKEN_CLIP_P,
    EmptyHotwordScorer,
    LanguageModel,
)

# UP:pyctcdecode/alphabet.py
UNK_TOKEN = "⁇"
UNK_TOKEN_PTN = re.compile(r"^[<\[]unk[>\]]$", flags=re.IGNORECASE)
BLANK_TOKEN_PTN = re.compile(r"^[<\[]pad[>\]]$", flags=re.IGNORECASE)
BPE_TOKEN = "▁"
UNK_BPE_TOKEN = "▁⁇▁"

NULL_FRAMES = (-1, -1)
EMPTY_START_BEAM = ("", "", "", None, [], NULL_FRAMES, 0.0)


def normalize_regular_alphabet(labels: list[str]) -> list[str]:
    """UP:pyctcdecode ``alphabet._normalize_regular_alphabet``."""
    normalized = labels[:]
    if "|" in normalized and " " not in normalized:
        normalized = [" " if c == "|" else c for c in normalized]
    for n, label in enumerate(normalized):
        if BLANK_TOKEN_PTN.match(label):
            normalized[n] = ""
    if "_" in normalized and "" not in normalized:
        normalized = ["" if c == "_" else c for c in normalized]
    if "" not in normalized:
        normalized.append("")
    for n, label in enumerate(normalized):
        if UNK_TOKEN_PTN.match(label):
            normalized[n] = UNK_TOKEN
    return normalized


def is_bpe_alphabet(labels: list[str]) -> bool:
    return any(s.startswith("##") for s in labels) or any(s.startswith(BPE_TOKEN) for s in labels)


def _log_softmax(x: np.ndarray, axis: int) -> np.ndarray:
    x_max = np.amax(x, axis=axis, keepdims=True)
    if x_max.ndim > 0:
        x_max[~np.isfinite(x_max)] = 0
    elif not np.isfinite(x_max):
        x_max = 0
    tmp = x - x_max
    exp_tmp = np.exp(tmp)
    with np.errstate(divide="ignore"):
        s = np.sum(exp_tmp, axis=axis, keepdims=True)
        out = np.log(s)
    return tmp - out


def prepare_logprobs(logits: np.ndarray) -> np.ndarray:
    """Input normalisation of ``decode_beams``.

    Returns float32-valued log-probs. Under the reference's numpy 1.26.4 the
    clipped array stays float32; under numpy 2 the clip against a float64 scalar
    promotes -- either way the values are float32-rounded and every later
    sum/compare is float64, which is what is forced here.
    """
    if math.isclose(float(logits.sum(axis=1).mean()), 1):
        lp = np.log(np.clip(logits, MIN_TOKEN_CLIP_P, 1))
        return lp.astype(np.float32) if lp.dtype != np.float32 else lp
    lo = np.float32(np.log(MIN_TOKEN_CLIP_P)) if logits.dtype == np.float32 else np.log(MIN_TOKEN_CLIP_P)
    return np.clip(_log_softmax(logits, axis=1), lo, logits.dtype.type(0))


def _merge_tokens(token_1: str, token_2: str) -> str:
    if len(token_2) == 0:
        return token_1
    if len(token_1) == 0:
        return token_2
    return token_1 + " " + token_2


def _sum_log_scores(s1: float, s2: float) -> float:
    if s1 >= s2:
        return s1 + math.log(1 + math.exp(s2 - s1))
    return s2 + math.log(1 + math.exp(s1 - s2))


def _merge_beams(beams):
    beam_dict = {}
    for text, next_word, word_part, last_char, text_frames, part_frames, logit_score in beams:
        new_text = _merge_tokens(text, next_word)
        hash_idx = (new_text, word_part, last_char)
        if hash_idx not in beam_dict:
            beam_dict[hash_idx] = (
                text, next_word, word_part, last_char, text_frames, part_frames, logit_score,
            )
        else:
            beam_dict[hash_idx] = (
                text, next_word, word_part, last_char, text_frames, part_frames,
                _sum_log_scores(beam_dict[hash_idx][-1], logit_score),
            )
    return list(beam_dict.values())


def _sort_and_trim_beams(beams, beam_width: int):
    return heapq.nlargest(beam_width, beams, key=lambda x: x[-1])


def _prune_history(beams, lm_order: int):
    min_n_history = max(1, lm_order - 1)
    seen_hashes = set()
    filtered_beams = []
    for text, next_word, word_part, last_char, text_frames, part_frames, logit_score, _ in beams:
        hash_idx = (tuple(text.split()[-min_n_history:]), word_part, last_char)
        if hash_idx not in seen_hashes:
            filtered_beams.append(
                (text, next_word, word_part, last_char, text_frames, part_frames, logit_score)
            )
            seen_hashes.add(hash_idx)
    return filtered_beams


class Alphabet:
    def __init__(self, labels: list[str], is_bpe: bool) -> None:
        self._labels = labels
        self._is_bpe = is_bpe

    @property
    def is_bpe(self) -> bool:
        return self._is_bpe

    @property
    def labels(self) -> list[str]:
        return self._labels[:]

    @classmethod
    def build_alphabet(cls, labels: list[str]) -> "Alphabet":
        if is_bpe_alphabet(labels):
            raise NotImplementedError("BPE alphabets are unreachable for CoRal (SURVEY A5)")
        return cls(normalize_regular_alphabet(labels), False)


class BeamSearchDecoderCTC:
    """Oracle decoder. ``stats`` accumulates the work counters SURVEY 8d defines."""

    def __init__(self, alphabet: Alphabet, language_model: LanguageModel | None = None) -> None:
        self._alphabet = alphabet
        self._idx2vocab = {n: c for n, c in enumerate(self._alphabet.labels)}
        self._language_model = language_model
        self.stats = {"frames": 0, "extensions": 0, "n_score": 0, "n_partial": 0, "probes": 0}

    def reset_params(self, alpha=None, beta=None, unk_score_offset=None, lm_score_boundary=None):
        lm = self._language_model
        if lm is None:
            return
        if alpha is not None:
            lm.alpha = alpha
        if beta is not None:
            lm.beta = beta
        if unk_score_offset is not None:
            lm.unk_score_offset = unk_score_offset
        if lm_score_boundary is not None:
            lm.score_boundary = lm_score_boundary

    def _check_logits_dimension(self, logits: np.ndarray) -> None:
        if len(logits.shape) != 2:
            raise ValueError(
                "Input logits have %s dimensions, but need 2: (time, vocabulary)" % len(logits.shape)
            )
        if logits.shape[-1] != len(self._idx2vocab):
            raise ValueError(
                "Input logits shape is %s, but vocabulary is size %s. "
                "Need logits of shape: (time, vocabulary)" % (logits.shape, len(self._idx2vocab))
            )

    def _get_lm_beams(self, beams, hotword_scorer, cached_lm_scores, cached_partial_token_scores,
                      is_eos: bool = False):
        language_model = self._language_model
        if language_model is None:
            new_beams = []
            for text, next_word, word_part, last_char, frame_list, frames, logit_score in beams:
                new_text = _merge_tokens(text, next_word)
                lm_hw_score = (
                    logit_score
                    + hotword_scorer.score(new_text)
                    + hotword_scorer.score_partial_token(word_part)
                )
                new_beams.append(
                    (new_text, "", word_part, last_char, frame_list, frames, logit_score, lm_hw_score)
                )
            return new_beams

        new_beams = []
        for text, next_word, word_part, last_char, frame_list, frames, logit_score in beams:
            new_text = _merge_tokens(text, next_word)
            if (new_text, is_eos) not in cached_lm_scores:
                _, prev_raw_lm_score, start_state = cached_lm_scores[(text, False)]
                score, end_state = language_model.score(start_state, next_word, is_last_word=is_eos)
                raw_lm_score = prev_raw_lm_score + score
                lm_hw_score = raw_lm_score + hotword_scorer.score(new_text)
                cached_lm_scores[(new_text, is_eos)] = (lm_hw_score, raw_lm_score, end_state)
            lm_score, _, _ = cached_lm_scores[(new_text, is_eos)]
            if len(word_part) > 0:
                if word_part not in cached_partial_token_scores:
                    if word_part in hotword_scorer:
                        cached_partial_token_scores[word_part] = hotword_scorer.score_partial_token(word_part)
                    else:
                        cached_partial_token_scores[word_part] = language_model.score_partial_token(word_part)
                lm_score += cached_partial_token_scores[word_part]
            new_beams.append(
                (new_text, "", word_part, last_char, frame_list, frames, logit_score,
                 logit_score + lm_score)
            )
        return new_beams

    def _decode_logits(self, logits, beam_width, beam_prune_logp, token_min_logp, prune_history,
                       hotword_scorer, lm_start_state=None):
        language_model = self._language_model
        if lm_start_state is None and language_model is not None:
            cached_lm_scores = {("", False): (0.0, 0.0, language_model.get_start_state())}
        else:
            cached_lm_scores = {("", False): (0.0, 0.0, lm_start_state)}
        cached_p_lm_scores: dict[str, float] = {}
        beams = [EMPTY_START_BEAM]
        # token_min_logp is compared in the log-prob dtype (float32), SURVEY A5 step 4
        thr = logits.dtype.type(token_min_logp)
        for frame_idx, logit_col in enumerate(logits):
            max_idx = int(logit_col.argmax())
            idx_list = sorted(set(int(i) for i in np.where(logit_col >= thr)[0]) | {max_idx})
            new_beams = []
            for idx_char in idx_list:
                p_char = float(logit_col[idx_char])  # float32 value, float64 arithmetic
                char = self._idx2vocab[idx_char]
                for text, next_word, word_part, last_char, text_frames, part_frames, logit_score in beams:
                    if char == "" or last_char == char:
                        if char == "":
                            new_end_frame = part_frames[0]
                        else:
                            new_end_frame = frame_idx + 1
                        new_part_frames = (
                            part_frames if char == "" else (part_frames[0], new_end_frame)
                        )
                        new_beams.append(
                            (text, next_word, word_part, char, text_frames, new_part_frames,
                             logit_score + p_char)
                        )
                    elif char == " ":
                        new_frame_list = (
                            text_frames if word_part == "" else text_frames + [part_frames]
                        )
                        new_beams.append(
                            (text, word_part, "", char, new_frame_list, NULL_FRAMES,
                             logit_score + p_char)
                        )
                    else:
                        new_part_frames = (
                            (frame_idx, frame_idx + 1)
                            if part_frames[0] < 0
                            else (part_frames[0], frame_idx + 1)
                        )
                        new_beams.append(
                            (text, next_word, word_part + char, char, text_frames, new_part_frames,
                             logit_score + p_char)
                        )
            self.stats["frames"] += 1
            self.stats["extensions"] += len(new_beams)
            new_beams = _merge_beams(new_beams)
            scored_beams = self._get_lm_beams(new_beams, hotword_scorer, cached_lm_scores, cached_p_lm_scores)
            max_score = max([b[-1] for b in scored_beams])
            scored_beams = [b for b in scored_beams if b[-1] >= max_score + beam_prune_logp]
            trimmed_beams = _sort_and_trim_beams(scored_beams, beam_width)
            if prune_history:
                lm_order = 1 if language_model is None else language_model.order
                beams = _prune_history(trimmed_beams, lm_order=lm_order)
            else:
                beams = [b[:-1] for b in trimmed_beams]

        new_beams = []
        for text, _, word_part, _, frame_list, frames, logit_score in beams:
            new_token_times = frame_list if word_part == "" else frame_list + [frames]
            new_beams.append((text, word_part, "", None, new_token_times, (-1, -1), logit_score))
        new_beams = _merge_beams(new_beams)
        scored_beams = self._get_lm_beams(
            new_beams, hotword_scorer, cached_lm_scores, cached_p_lm_scores, is_eos=True
        )
        max_score = max([b[-1] for b in scored_beams])
        scored_beams = [b for b in scored_beams if b[-1] >= max_score + beam_prune_logp]
        trimmed_beams = _sort_and_trim_beams(scored_beams, beam_width)
        output_beams = [
            (
                " ".join(text.split()),
                cached_lm_scores[(text, True)][-1] if (text, True) in cached_lm_scores else None,
                list(zip(text.split(), text_frames)),
                logit_score,
                combined_score,
            )
            for text, _, _, _, text_frames, _, logit_score, combined_score in trimmed_beams
        ]
        return output_beams

    def decode_beams(self, logits, beam_width=DEFAULT_BEAM_WIDTH, beam_prune_logp=DEFAULT_PRUNE_LOGP,
                     token_min_logp=DEFAULT_MIN_TOKEN_LOGP, prune_history=False, hotwords=None,
                     hotword_weight=DEFAULT_HOTWORD_WEIGHT, lm_start_state=None):
        self._check_logits_dimension(logits)
        if hotwords:
            raise NotImplementedError("hotwords are never passed by CoRal (SURVEY A9)")
        hotword_scorer = EmptyHotwordScorer()
        logprobs = prepare_logprobs(logits)
        lm = self._language_model
        if lm is not None:
            c0 = (lm.n_score_calls, lm.n_partial_calls, lm.n_score_probes)
        out = self._decode_logits(logprobs, beam_width, beam_prune_logp, token_min_logp,
                                  prune_history, hotword_scorer, lm_start_state)
        if lm is not None:
            self.stats["n_score"] += lm.n_score_calls - c0[0]
            self.stats["n_partial"] += lm.n_partial_calls - c0[1]
            self.stats["probes"] += lm.n_score_probes - c0[2]
        return out

    def decode(self, logits, **kwargs) -> str:
        return self.decode_beams(logits, **kwargs)[0][0]

    def decode_beams_batch(self, pool, logits_list, **kwargs):
        """MP-safe 4-tuples (the ``kenlm.State`` is dropped), sequentially."""
        out = []
        for logits in logits_list:
            beams = self.decode_beams(logits, **kwargs)
            out.append([(t, frames, ls, lms) for t, _, frames, ls, lms in beams])
        return out

    def decode_batch(self, pool, logits_list, **kwargs) -> list[str]:
        return [self.decode(logits, **kwargs) for logits in logits_list]


def build_ctcdecoder(labels, kenlm_model_path=None, unigrams=None, alpha=DEFAULT_ALPHA,
                     beta=DEFAULT_BETA, unk_score_offset=DEFAULT_UNK_LOGP_OFFSET,
                     lm_score_boundary=DEFAULT_SCORE_LM_BOUNDARY) -> BeamSearchDecoderCTC:
    """UP:pyctcdecode ``decoder.build_ctcdecoder``, ARPA only."""
    alphabet = Alphabet.build_alphabet(list(labels))
    if kenlm_model_path is None:
        return BeamSearchDecoderCTC(alphabet, None)
    kenlm_model = ArpaModel.load(kenlm_model_path)
    if str(kenlm_model_path).endswith(".arpa") and unigrams is None:
        unigrams = load_unigram_set_from_arpa(kenlm_model_path)
    lm = LanguageModel(kenlm_model, unigrams, alpha=alpha, beta=beta,
                       unk_score_offset=unk_score_offset, score_boundary=lm_score_boundary)
    return BeamSearchDecoderCTC(alphabet, lm)
