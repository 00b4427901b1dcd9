"""jiwer 4.0.0 / rapidfuzz 3.14.3 edit counts and CoRal's CER/WER (test infrastructure only).

Restates, per SURVEY.md section 8 A1/A2/A11/A12/A13:

* R:src/coral/metrics.py:8-33 (``cer``) and :36-61 (``wer``);
* R:src/coral/evaluate.py:161-216 (``get_score_df``);
* UP:jiwer ``process.py`` / ``transforms.py`` default transforms and ``_word2char``;
* UP:rapidfuzz-cpp ``distance/Levenshtein_impl.hpp``: ``remove_common_affix``, the
  Hyyro 2003 bit-parallel matrix (VP/VN rows recorded per character of the second
  sequence) and ``recover_alignment``'s backtrace preference.

Parity unpinned -- see ``oracle/__init__.py``. The DP here is the plain O(nm)
matrix; ``VP`` bit (row j, col i) == ``D[i][j] == D[i-1][j] + 1`` and ``VN`` bit
== ``D[i][j] == D[i-1][j] - 1`` where ``D[i][j]`` is the distance between
``s1[:i]`` and ``s2[:j]``.
"""

from __future__ import annotations

import itertools as it
import re

import numpy as np

_MULTI_SPACE = re.compile(r"\s\s+")


# --------------------------------------------------------------------------- jiwer
def words_transform(s: str) -> list[str]:
    """``wer_default``: RemoveMultipleSpaces -> Strip -> ReduceToListOfListOfWords."""
    s = _MULTI_SPACE.sub(" ", s)
    s = s.strip()
    return [w for w in s.split(" ") if len(w) >= 1]


def chars_transform(s: str) -> list[str]:
    """``cer_default``: Strip -> ReduceToListOfListOfChars."""
    return list(s.strip())


# jiwer >= 3.1 (the reference pins 4.0.0, R:uv.lock:1204-1205) accepts an empty reference and
# scores it as all insertions; 3.0.x raised ValueError here. Restated from the upstream change
# log -- unpinned like the rest of this module; "raise" restores the old behaviour.
EMPTY_REFERENCE = "allow"


def _check_reference(reference: str, seq: list) -> None:
    if EMPTY_REFERENCE == "raise" and (len(reference) == 0 or len(seq) == 0):
        raise ValueError("one or more references are empty strings")


# ----------------------------------------------------------------------- rapidfuzz
def remove_common_affix(s1, s2):
    n1, n2 = len(s1), len(s2)
    p = 0
    while p < n1 and p < n2 and s1[p] == s2[p]:
        p += 1
    s = 0
    while s < n1 - p and s < n2 - p and s1[n1 - 1 - s] == s2[n2 - 1 - s]:
        s += 1
    return s1[p : n1 - s], s2[p : n2 - s]


def editops_counts(s1, s2) -> tuple[int, int, int]:
    """(substitutions, deletions, insertions) of ``Levenshtein.editops(s1, s2)``."""
    s1, s2 = remove_common_affix(list(s1), list(s2))
    n1, n2 = len(s1), len(s2)
    if n1 == 0 or n2 == 0:
        return 0, n1, n2
    D = np.zeros((n1 + 1, n2 + 1), dtype=np.int64)
    D[:, 0] = np.arange(n1 + 1)
    D[0, :] = np.arange(n2 + 1)
    for j in range(1, n2 + 1):
        c = s2[j - 1]
        for i in range(1, n1 + 1):
            cost = 0 if s1[i - 1] == c else 1
            D[i, j] = min(D[i - 1, j] + 1, D[i, j - 1] + 1, D[i - 1, j - 1] + cost)
    S = Dl = I = 0
    col, row = n1, n2
    while row and col:
        if D[col, row] == D[col - 1, row] + 1:  # VP bit (row-1, col-1)
            Dl += 1
            col -= 1
        else:
            row -= 1
            if row and D[col, row] == D[col - 1, row] - 1:  # VN bit (row-1, col-1)
                I += 1
            else:
                col -= 1
                if s1[col] != s2[row]:
                    S += 1
    Dl += col
    I += row
    return S, Dl, I


def editops_counts_fast(s1, s2) -> tuple[int, int, int]:
    """Same result as :func:`editops_counts` with numpy rows (for larger cases)."""
    s1, s2 = remove_common_affix(list(s1), list(s2))
    n1, n2 = len(s1), len(s2)
    if n1 == 0 or n2 == 0:
        return 0, n1, n2
    # intern tokens to ints for vector compares
    table: dict = {}
    a = np.asarray([table.setdefault(x, len(table)) for x in s1], dtype=np.int64)
    b = [table.setdefault(x, len(table)) for x in s2]
    D = np.empty((n2 + 1, n1 + 1), dtype=np.int64)  # D[j][i]
    D[0] = np.arange(n1 + 1)
    idx = np.arange(n1 + 1)
    for j in range(1, n2 + 1):
        prev = D[j - 1]
        cost = (a != b[j - 1]).astype(np.int64)
        # candidates without the horizontal (i-1 -> i) dependency
        cand = np.empty(n1 + 1, dtype=np.int64)
        cand[0] = j
        cand[1:] = np.minimum(prev[1:] + 1, prev[:-1] + cost)
        # resolve D[j][i] = min(cand[i], D[j][i-1] + 1) with a running min
        D[j] = np.minimum.accumulate(cand - idx) + idx
    S = Dl = I = 0
    col, row = n1, n2
    while row and col:
        if D[row, col] == D[row, col - 1] + 1:
            Dl += 1
            col -= 1
        else:
            row -= 1
            if row and D[row, col] == D[row, col - 1] - 1:
                I += 1
            else:
                col -= 1
                if s1[col] != s2[row]:
                    S += 1
    Dl += col
    I += row
    return S, Dl, I


# ------------------------------------------------------------------ per-pair counts
def char_counts(reference: str, hypothesis: str) -> tuple[int, int, int, int]:
    """``jiwer.process_characters(reference, hypothesis)`` -> (S, D, I, H)."""
    ref = chars_transform(reference)
    _check_reference(reference, ref)
    hyp = chars_transform(hypothesis)
    S, Dl, I = editops_counts_fast(ref, hyp)
    return S, Dl, I, len(ref) - (S + Dl)


def word_counts(reference: str, hypothesis: str) -> tuple[int, int, int, int]:
    """``jiwer.process_words(reference, hypothesis)`` -> (S, D, I, H)."""
    ref = words_transform(reference)
    _check_reference(reference, ref)
    hyp = words_transform(hypothesis)
    S, Dl, I = editops_counts_fast(ref, hyp)
    return S, Dl, I, len(ref) - (S + Dl)


# ------------------------------------------------------------- R:src/coral/metrics.py
def _rate(predictions, labels, normalise: bool, counts) -> float:
    incorrect = 0
    total = 0
    for prediction, label in zip(predictions, labels):
        S, Dl, I, H = counts(label, prediction)
        incorrect += S + Dl + I
        total += S + Dl + H
        if normalise:
            total += I
    return incorrect / total


def cer(predictions, labels, normalise: bool = True) -> float:
    return _rate(predictions, labels, normalise, char_counts)


def wer(predictions, labels, normalise: bool = True) -> float:
    return _rate(predictions, labels, normalise, word_counts)


def per_sample_rates(predictions, labels, normalise: bool = True, kind: str = "cer") -> list[float]:
    """Each sample scored alone with the same formula (what
    R:src/coral/validation.py:149-158 needs as the per-row ``asr_cer``)."""
    f = char_counts if kind == "cer" else word_counts
    return [_rate([p], [l], normalise, f) for p, l in zip(predictions, labels)]


# ----------------------------------------------------- R:src/coral/evaluate.py:161-216
def get_score_records(rows: list[dict], categories: list[str]) -> list[dict]:
    """``get_score_df`` on a list of row dicts (keys: categories, prediction, text)."""
    uniques = []
    for cat in categories:
        seen: list = []
        for r in rows:
            if r[cat] not in seen:
                seen.append(r[cat])
        uniques.append(seen + [None])
    records = []
    for combination in it.product(*uniques):
        filtered = rows
        skip = False
        for key, value in zip(categories, combination):
            if value is None:
                continue
            new = [r for r in filtered if r[key] == value]
            if len(new) == len(filtered) or len(new) == 0:
                skip = True
            filtered = new
        if skip:
            continue
        preds = [r["prediction"] for r in filtered]
        labs = [r["text"] for r in filtered]
        records.append(dict(zip(categories, combination)) | dict(cer=cer(preds, labs), wer=wer(preds, labs)))
    return records
