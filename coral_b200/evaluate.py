"""Drop-in for ``coral.evaluate.get_score_df`` (R:src/coral/evaluate.py:161-216).

The reference re-aligns every utterance once per demographic combination (about
130-180 subsets). Here every (prediction, label) pair is aligned ONCE on
the GPU (chars and words), and each combination is an integer sum over the rows it
selects -- the same numbers, because CER/WER of a subset only depend on the per-pair
(S, D, I, H). The combinations, their order, the skip rule (:187-192) and the record
layout are the reference's.
"""

from __future__ import annotations

import itertools as it
import logging

import numpy as np
import pandas as pd

from .metrics import _rate_from_counts

logger = logging.getLogger(__name__)


def group_masks(df: pd.DataFrame, categories: list[str]):
    """Yield ``(combination, boolean row mask)`` exactly as the reference enumerates them."""
    unique_category_values = [df[category].unique().tolist() + [None] for category in categories]
    cols = {c: df[c].to_numpy() for c in categories}
    n = len(df)
    for combination in it.product(*unique_category_values):
        mask = np.ones(n, dtype=bool)
        skip_combination = False
        for key, value in zip(categories, combination):
            if value is None:
                continue
            new_mask = mask & (cols[key] == value)
            if new_mask.sum() == mask.sum() or new_mask.sum() == 0:
                skip_combination = True
            mask = new_mask
        if skip_combination:
            continue
        yield combination, mask


def score_records(df: pd.DataFrame, categories: list[str], char_counts: np.ndarray, word_counts: np.ndarray):
    records = []
    for combination, mask in group_masks(df, categories):
        named_combination = dict(zip(categories, combination))
        score_dict = dict(
            cer=_rate_from_counts(char_counts[mask], True),
            wer=_rate_from_counts(word_counts[mask], True),
        )
        records.append(named_combination | score_dict)
        combination_str = ", ".join(f"{k}={v}" for k, v in named_combination.items() if v is not None)
        if combination_str == "":
            combination_str = "entire dataset"
        score_str = ", ".join(f"{k.upper()} = {v:.1%}" for k, v in score_dict.items())
        logger.info(f"Scores for {combination_str}: {score_str}")
    return records


def get_score_df(df: pd.DataFrame, categories: list[str]) -> pd.DataFrame:
    """``df`` needs the category columns plus ``prediction`` and ``text`` (the label)."""
    from .metrics import _pair_counts

    preds = df.prediction.tolist()
    labs = df.text.tolist()
    both = _pair_counts(preds, labs, ("chars", "words"))  # aligned once; the groups only sum integers
    char_counts, word_counts = both["chars"], both["words"]
    return pd.DataFrame.from_records(data=score_records(df, categories, char_counts, word_counts))
