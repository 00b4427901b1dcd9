"""The metric step of ``coral.validation.add_validations`` (R:src/coral/validation.py:136-159).

The reference computes the aggregate ``cer`` / ``wer`` (:137-140), then adds
``asr_cer`` / ``asr_wer`` columns (:149-152) and keeps the rows with
``asr_cer < max_cer`` (:156-158). As written it passes the aggregate float where a
per-row column is needed (SURVEY.md section 3.2 documents the defect); the intended contract,
visible in the published dataset, is a per-sample score. ``validation_scores`` returns
both: the aggregates with the reference's exact formula and the per-sample columns
(each sample scored alone with the same formula), plus the keep-mask.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .metrics import _rate_from_counts, per_sample_rates


@dataclass
class ValidationScores:
    cer: float                 # aggregate, R:src/coral/validation.py:138
    wer: float                 # aggregate, R:src/coral/validation.py:139
    asr_cer: np.ndarray        # per sample, float64 [n]
    asr_wer: np.ndarray        # per sample, float64 [n]
    keep: np.ndarray           # asr_cer < max_cer, bool [n]
    char_counts: np.ndarray    # [n, 4] S, D, I, H
    word_counts: np.ndarray    # [n, 4]


def validation_scores(predictions, labels, max_cer: float = 0.6, normalise: bool = True) -> ValidationScores:
    from .metrics import _as_lists, _pair_counts

    predictions, labels = _as_lists(predictions if isinstance(predictions, list) else list(predictions),
                                    labels if isinstance(labels, list) else list(labels))
    both = _pair_counts(predictions, labels, ("chars", "words"))  # one marshalling, both kernels, one read-back
    cc, wc = both["chars"], both["words"]
    asr_cer = per_sample_rates(cc, normalise)
    asr_wer = per_sample_rates(wc, normalise)
    return ValidationScores(
        cer=_rate_from_counts(cc, normalise), wer=_rate_from_counts(wc, normalise),
        asr_cer=asr_cer, asr_wer=asr_wer, keep=asr_cer < max_cer, char_counts=cc, word_counts=wc,
    )
