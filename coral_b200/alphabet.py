"""pyctcdecode-compatible alphabet and constants (host-side mirror of the reference interface).

Mirrors UP:pyctcdecode 0.5.0 ``alphabet.py`` / ``constants.py`` as specified in SURVEY.md
section 8 A5/A6; the names below are the ones HF imports
(HF:models/wav2vec2_with_lm/processing_wav2vec2_with_lm.py:187, :351-356).
"""

from __future__ import annotations

import json
import logging
import re

import numpy as np

logger = logging.getLogger(__name__)

# ---- pyctcdecode.constants
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

# ---- pyctcdecode.alphabet
UNK_TOKEN = "⁇"
UNK_TOKEN_PTN = re.compile(r"^[<\[]unk[>\]]$", flags=re.IGNORECASE)
BLANK_TOKEN_PTN = re.compile(r"^[<\[]pad[>\]]$", flags=re.IGNORECASE)
BPE_TOKEN = "▁"
UNK_BPE_TOKEN = "▁⁇▁"


def _normalize_regular_alphabet(labels: list[str]) -> list[str]:
    normalized = labels[:]
    if "|" in normalized and " " not in normalized:
        logger.info("Found '|' in vocabulary but not ' ', doing substitution.")
        normalized = [" " if c == "|" else c for c in normalized]
    for n, label in enumerate(normalized):
        if BLANK_TOKEN_PTN.match(label):
            normalized[n] = ""
    if "_" in normalized and "" not in normalized:
        logger.info("Found '_' in vocabulary but not '', doing substitution.")
        normalized = ["" if c == "_" else c for c in normalized]
    if "" not in normalized:
        logger.info("Blank token not found in vocabulary, appending it.")
        normalized.append("")
    for n, label in enumerate(normalized):
        if UNK_TOKEN_PTN.match(label):
            normalized[n] = UNK_TOKEN
    if any(len(c) > 1 for c in normalized):
        logger.warning("Found entries of length > 1 in alphabet. This is unusual unless style is BPE.")
    return normalized


def _is_bpe(labels: list[str]) -> bool:
    return any(s.startswith("##") for s in labels) or any(s.startswith(BPE_TOKEN) for s in labels)


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

    def dumps(self) -> str:
        return json.dumps({"labels": self._labels, "is_bpe": self._is_bpe}, indent=2)

    @classmethod
    def loads(cls, json_data: str) -> "Alphabet":
        data = json.loads(json_data)
        return cls(data["labels"], data["is_bpe"])

    @classmethod
    def build_alphabet(cls, labels: list[str]) -> "Alphabet":
        labels = list(labels)
        if _is_bpe(labels):
            raise NotImplementedError(
                "BPE alphabets are not supported: CoRal's vocabulary is character-level "
                "(R:src/coral/wav2vec2.py:318-322; SURVEY.md section 8 A5)"
            )
        return cls(_normalize_regular_alphabet(labels), False)
