"""pyctcdecode ``LanguageModel`` restated over ``oracle.arpa`` (test infrastructure only).

Follows UP:pyctcdecode 0.5.0 ``language_model.py`` (``LanguageModel.__init__``,
``score``, ``score_partial_token``, ``get_start_state``, ``_get_raw_end_score``,
``HotwordScorer``) as specified in SURVEY.md section 8 A7/A9; the reference reaches it
through ``build_ctcdecoder`` (R:src/coral/ngram.py:341-343) and through
``Wav2Vec2ProcessorWithLM`` (HF:models/wav2vec2_with_lm/processing_wav2vec2_with_lm.py:
160-183 sets ``alpha/beta/unk_score_offset/score_boundary`` on this object).
Parity unpinned -- see ``oracle/__init__.py``.
"""

from __future__ import annotations

import numpy as np

from .arpa import ArpaModel, State

# UP:pyctcdecode/constants.py
DEFAULT_ALPHA = 0.5
DEFAULT_BETA = 1.5
DEFAULT_UNK_LOGP_OFFSET = -10.0
DEFAULT_BEAM_WIDTH = 100
DEFAULT_HOTWORD_WEIGHT = 10.0
DEFAULT_PRUNE_LOGP = -10.0
DEFAULT_PRUNE_BEAMS = False
DEFAULT_MIN_TOKEN_LOGP = -5.0
DEFAULT_SCORE_LM_BOUNDARY = True
AVG_TOKEN_LEN = 6
MIN_TOKEN_CLIP_P = 1e-15
LOG_BASE_CHANGE_FACTOR = 1.0 / np.log10(np.e)


class EmptyHotwordScorer:
    """The scorer pyctcdecode builds for ``hotwords=None``: all zeros."""

    def __contains__(self, item: str) -> bool:
        return False

    def score(self, text: str) -> float:
        return 0.0

    def score_partial_token(self, token: str) -> float:
        return 0.0


class CharTrieSet:
    """``pygtrie.CharTrie.fromkeys(unigrams)`` reduced to what is used: ``has_node``."""

    def __init__(self, words):
        self._nodes: set[str] = set()
        for w in words:
            for i in range(1, len(w) + 1):
                self._nodes.add(w[:i])

    def has_node(self, key: str) -> bool:
        return key in self._nodes


class LanguageModel:
    """UP:pyctcdecode ``LanguageModel`` over an :class:`ArpaModel`."""

    def __init__(
        self,
        kenlm_model: ArpaModel,
        unigrams=None,
        alpha: float = DEFAULT_ALPHA,
        beta: float = DEFAULT_BETA,
        unk_score_offset: float = DEFAULT_UNK_LOGP_OFFSET,
        score_boundary: bool = DEFAULT_SCORE_LM_BOUNDARY,
    ) -> None:
        self._kenlm_model = kenlm_model
        if unigrams is None:
            unigram_set: set[str] = set()
            char_trie = None
        else:
            unigram_set = set(t for t in set(unigrams) if t in self._kenlm_model)
            char_trie = CharTrieSet(unigram_set)
        self._unigram_set = unigram_set
        self._char_trie = char_trie
        self.alpha = alpha
        self.beta = beta
        self.unk_score_offset = unk_score_offset
        self.score_boundary = score_boundary
        # counters for SURVEY 8d's algorithmic-byte figure
        self.n_score_calls = 0
        self.n_score_probes = 0
        self.n_partial_calls = 0

    @property
    def order(self) -> int:
        return self._kenlm_model.order

    def get_start_state(self) -> State:
        if self.score_boundary:
            return self._kenlm_model.begin_sentence_state()
        return self._kenlm_model.null_context_state()

    def _get_raw_end_score(self, start_state: State) -> float:
        if self.score_boundary:
            self.n_score_probes += self._kenlm_model.probes(
                start_state, self._kenlm_model.index("</s>")
            )
            return self._kenlm_model.base_score(start_state, "</s>")[0]
        return 0.0

    def score_partial_token(self, partial_token: str) -> float:
        self.n_partial_calls += 1
        if self._char_trie is None:
            is_oov = 1.0
        else:
            is_oov = int(self._char_trie.has_node(partial_token) == 0)
        unk_score = self.unk_score_offset * is_oov
        if len(partial_token) > AVG_TOKEN_LEN:
            unk_score = unk_score * len(partial_token) / AVG_TOKEN_LEN
        return unk_score

    def score(self, prev_state: State, word: str, is_last_word: bool = False):
        self.n_score_calls += 1
        self.n_score_probes += self._kenlm_model.probes(
            prev_state, self._kenlm_model.index(word)
        )
        lm_score, end_state = self._kenlm_model.base_score(prev_state, word)
        if (
            len(self._unigram_set) > 0
            and word not in self._unigram_set
            or word not in self._kenlm_model
        ):
            lm_score += self.unk_score_offset
        if is_last_word:
            lm_score = lm_score + self._get_raw_end_score(end_state)
        lm_score = self.alpha * lm_score * LOG_BASE_CHANGE_FACTOR + self.beta
        return float(lm_score), end_state
