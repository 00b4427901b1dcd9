"""Pins that do not come from this repo's oracle ("what's weak" 1):

* ``torchaudio.functional.edit_distance`` (installed, 2.11; an independent Levenshtein written by
  other people) against ``S + D + I`` -- of the oracle on the CPU and of the CUDA kernel on the GPU --
  on large random sets that include > 64- and > 128-symbol strings (rapidfuzz's block path) and
  word sequences;
* ``tests/golden/published_known_answers.json``: outputs of the real jiwer / rapidfuzz as printed
  in their own documentation (one of them decides the S/D/I tie-break) and hand-walked DP matrices.
"""

from __future__ import annotations

import json
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "published_known_answers.json")
ALPHA = list("abcdefghijklmnopqrstuvwxyzæøå")


def _pairs(rng, n, lo, hi, p_edit=0.15, alphabet=ALPHA, spaces=True):
    """Reference = random words; hypothesis = the reference with i.i.d. edits (so distances are
    non-trivial) or, for one pair in eight, an unrelated string."""
    refs, hyps = [], []
    letters = np.array(alphabet + ([" "] * 6 if spaces else []))
    for k in range(n):
        L = int(rng.integers(lo, hi + 1))
        ref = "".join(letters[rng.integers(0, len(letters), size=L)]).strip() or "a"
        if k % 8 == 7:
            hyp = "".join(letters[rng.integers(0, len(letters), size=int(rng.integers(0, hi + 1)))])
        else:
            out = []
            r = rng.random(len(ref))
            kind = rng.integers(0, 3, size=len(ref))
            pick = letters[rng.integers(0, len(letters), size=len(ref))]
            for ch, x, kd, pc in zip(ref, r, kind, pick):
                if x >= p_edit:
                    out.append(ch)
                elif kd == 0:
                    out.append(pc)
                elif kd == 2:
                    out.extend((pc, ch))
            hyp = "".join(out)
        refs.append(ref)
        hyps.append(hyp)
    return refs, hyps


def _torchaudio_distance(refs, hyps, words=False):
    import torchaudio.functional as F

    if words:
        from oracle.edit import words_transform

        return [F.edit_distance(words_transform(r), words_transform(h)) for r, h in zip(refs, hyps)]
    return [F.edit_distance(r.strip(), h.strip()) for r, h in zip(refs, hyps)]


def test_oracle_distance_equals_torchaudio(rng):
    from oracle import edit as oe

    for lo, hi, n in ((1, 40, 6000), (65, 130, 300), (129, 400, 60)):
        refs, hyps = _pairs(rng, n, lo, hi)
        ta = _torchaudio_distance(refs, hyps)
        for r, h, d in zip(refs, hyps, ta):
            S, D, I, H = oe.char_counts(r, h)
            assert S + D + I == d, (r, h)
            assert S + D + H == len(r.strip()) and I - D == len(h.strip()) - len(r.strip())
    refs, hyps = _pairs(rng, 2000, 10, 160)
    for r, h, d in zip(refs, hyps, _torchaudio_distance(refs, hyps, words=True)):
        S, D, I, H = oe.word_counts(r, h)
        assert S + D + I == d, (r, h)


def test_oracle_on_published_and_hand_walked_answers():
    from oracle import edit as oe

    data = json.load(open(GOLDEN, encoding="utf-8"))
    for c in data["chars"]:
        assert list(oe.char_counts(c["ref"], c["hyp"])) == c["counts"], c
        if "D" in c:  # the matrix written in the fixture is the plain DP matrix
            a, b = c["ref"], c["hyp"]
            D = np.zeros((len(a) + 1, len(b) + 1), dtype=int)
            D[:, 0] = np.arange(len(a) + 1)
            D[0, :] = np.arange(len(b) + 1)
            for i in range(1, len(a) + 1):
                for j in range(1, len(b) + 1):
                    D[i, j] = min(D[i - 1, j] + 1, D[i, j - 1] + 1, D[i - 1, j - 1] + (a[i - 1] != b[j - 1]))
            assert D.tolist() == c["D"], c
    for c in data["words"]:
        assert list(oe.word_counts(c["ref"], c["hyp"])) == c["counts"], c
    t = data["words_totals"]
    refs = [data["words"][k]["ref"] for k in t["pairs"]]
    hyps = [data["words"][k]["hyp"] for k in t["pairs"]]
    assert oe.wer(hyps, refs, normalise=False) == t["wer"]  # jiwer's wer
    assert oe.wer(hyps, refs, normalise=True) == t["mer"]   # CoRal's normalise=True is jiwer's mer


@pytest.mark.gpu
def test_kernel_distance_equals_torchaudio_100k(rng):
    """>= 100k random pairs through the CUDA kernel: S + D + I == torchaudio's distance, the count
    identities hold, and the long-string (off-chip work area) path is included."""
    import torch

    assert torch.cuda.is_available()
    from coral_b200.metrics import edit_counts

    total = 0
    for lo, hi, n in ((1, 40, 90_000), (30, 128, 9_000), (65, 200, 2_000), (129, 600, 150)):
        refs, hyps = _pairs(rng, n, lo, hi)
        cc = edit_counts(hyps, refs, "chars")
        ta = np.array(_torchaudio_distance(refs, hyps))
        assert np.array_equal(cc[:, 0] + cc[:, 1] + cc[:, 2], ta), (lo, hi)
        rl = np.array([len(r.strip()) for r in refs])
        hl = np.array([len(h.strip()) for h in hyps])
        assert np.array_equal(cc[:, 0] + cc[:, 1] + cc[:, 3], rl) and np.array_equal(cc[:, 2] - cc[:, 1], hl - rl)
        total += n
    refs, hyps = _pairs(rng, 6000, 10, 160)
    wc = edit_counts(hyps, refs, "words")
    assert np.array_equal(wc[:, 0] + wc[:, 1] + wc[:, 2], np.array(_torchaudio_distance(refs, hyps, words=True)))
    assert total + 6000 >= 100_000


@pytest.mark.gpu
def test_kernel_on_published_and_hand_walked_answers():
    from coral_b200 import metrics
    from coral_b200.metrics import edit_counts

    data = json.load(open(GOLDEN, encoding="utf-8"))
    cc = edit_counts([c["hyp"] for c in data["chars"]], [c["ref"] for c in data["chars"]], "chars")
    assert cc.tolist() == [c["counts"] for c in data["chars"]]
    wc = edit_counts([c["hyp"] for c in data["words"]], [c["ref"] for c in data["words"]], "words")
    assert wc.tolist() == [c["counts"] for c in data["words"]]
    t = data["words_totals"]
    refs = [data["words"][k]["ref"] for k in t["pairs"]]
    hyps = [data["words"][k]["hyp"] for k in t["pairs"]]
    assert metrics.wer(hyps, refs, normalise=False) == t["wer"]
    assert metrics.wer(hyps, refs, normalise=True) == t["mer"]
