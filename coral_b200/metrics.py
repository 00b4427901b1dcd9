"""CER / WER on the GPU: drop-in for ``coral.metrics`` (R:src/coral/metrics.py:8-61).

``cer`` / ``wer`` keep the reference's signatures and semantics: per pair the
(substitutions, deletions, insertions, hits) that ``jiwer.process_characters`` /
``process_words`` return, ``incorrect += S + D + I``, ``total += S + D + H (+ I if
normalise)``, result ``incorrect / total`` (true division of two Python ints;
``ZeroDivisionError`` when total is 0; ``ValueError`` on an empty reference, as jiwer
raises). All pairs are scored by ONE launch of ``coral_edit_counts``; only the integer
sums happen on the host. ``edit_counts`` exposes the per-pair counts, which is what
``get_score_df`` (align once, group many times) and the per-sample validation scores use.
"""

from __future__ import annotations

import collections.abc as c

import numpy as np

from . import _lib
from .textio import encode_utf32

MODE_TOKENS, MODE_CHARS, MODE_WORDS = 0, 1, 2

# What an EMPTY reference does. The reference pins jiwer 4.0.0 (R:uv.lock:1204-1205); jiwer accepts
# empty references since 3.1 and scores them as all-insertions (S = D = H = 0, I = len(hyp)), where
# 3.0.x raised ValueError("one or more references are empty strings"). jiwer is not installable in
# the build image, so this is restated from the upstream change log, not checked against the
# package (tools/pin_oracle.py records the real behaviour wherever jiwer imports). Default: what
# 4.0.0 does; CORAL_B200_EMPTY_REFERENCE=raise restores the 3.0.x error.
import os as _os

EMPTY_REFERENCE = _os.environ.get("CORAL_B200_EMPTY_REFERENCE", "allow")


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("coral_b200 needs a CUDA device: there is no CPU path")
    return torch


def edit_counts_device(ref_cps, ref_off, hyp_cps, hyp_off, n_pairs: int, mode: int, max_len: int, device=None):
    """Launch on already-resident buffers; returns device tensors (sdih int32 [n, 4], status int32 [n])."""
    torch = _torch()
    dev = ref_off.device if device is None else device
    sdih = torch.empty((max(n_pairs, 1), 4), dtype=torch.int32, device=dev)
    status = torch.empty(max(n_pairs, 1), dtype=torch.int32, device=dev)
    _lib.check(_lib.load().coral_edit_counts(
        ref_cps.data_ptr(), ref_off.data_ptr(), hyp_cps.data_ptr(), hyp_off.data_ptr(), n_pairs, mode, int(max_len),
        dev.index if dev.index is not None else torch.cuda.current_device(), sdih.data_ptr(), status.data_ptr(),
        _lib.stream_ptr(dev)))
    return sdih[:n_pairs], status[:n_pairs]


def edit_counts_spans_device(ref_cps, ref_beg, ref_end, hyp_cps, hyp_beg, hyp_end, n_pairs: int, mode: int,
                             max_len: int, out=None):
    """Like :func:`edit_counts_device` with explicit ``[begin, end)`` spans per string, e.g.
    hypotheses still sitting in the padded decoder output (begin = row * pitch)."""
    torch = _torch()
    dev = ref_beg.device
    if out is None:
        out = (torch.empty((max(n_pairs, 1), 4), dtype=torch.int32, device=dev),
               torch.empty(max(n_pairs, 1), dtype=torch.int32, device=dev))
    sdih, status = out
    _lib.check(_lib.load().coral_edit_counts_spans(
        ref_cps.data_ptr(), ref_beg.data_ptr(), ref_end.data_ptr(), hyp_cps.data_ptr(), hyp_beg.data_ptr(),
        hyp_end.data_ptr(), n_pairs, mode, int(max_len),
        dev.index if dev.index is not None else torch.cuda.current_device(), sdih.data_ptr(), status.data_ptr(),
        _lib.stream_ptr(dev)))
    return sdih[:n_pairs], status[:n_pairs]


# The reference calls cer() and then wer() on the same two lists (R:src/coral/validation.py:137-140,
# R:src/coral/evaluate.py:195-198). The first of the two marshals the strings once and queues BOTH
# kernels (characters and words); the second call finds its counts already computed. The memo
# holds one entry, keyed by the content fingerprints of the two lists, and lives on this object
# rather than in loose module globals.
class _PairMemo:
    def __init__(self):
        self.key = None
        self.counts = None

    def clear(self):
        self.key = None
        self.counts = None


_MEMO = _PairMemo()


def _fingerprint(strings) -> tuple:
    """Identity of a list's CONTENT. Transcripts fresh out of ``decode_batch`` carry a token (any
    in-place mutation drops it); everything else is fingerprinted by value: str hashes are cached
    by CPython, so this is ~0.1-0.6 ms for 8k strings."""
    held = getattr(strings, "_coral_dev", None)
    if held is not None:
        return held[3]
    return (len(strings), hash(tuple(strings)))


def _to_device(strings, dev):
    """(cps int32 tensor, offsets int64 tensor, max_len) of a list of str on ``dev``. Transcripts
    that came out of ``decode_batch`` are already there (``DecodedTexts._coral_dev``)."""
    torch = _torch()
    held = getattr(strings, "_coral_dev", None)
    if held is not None and held[0].device == dev:
        return held[0], held[1], held[2]
    try:
        cps, off = encode_utf32(strings)
    except TypeError as e:  # "".join refuses anything that is not a str
        raise TypeError("predictions and references must be strings") from e
    max_len = int(np.diff(off).max()) if len(off) > 1 else 0
    d_cps = torch.from_numpy(cps.view(np.int32)).to(dev, non_blocking=True)
    d_off = torch.from_numpy(off).to(dev, non_blocking=True)
    return d_cps, d_off, max_len


def _as_lists(predictions, labels):
    if isinstance(predictions, list) and isinstance(labels, list) and len(predictions) == len(labels):
        return predictions, labels
    pairs = list(zip(predictions, labels))  # the reference zips: the shorter iterable wins
    return [p for p, _ in pairs], [l for _, l in pairs]


def _pair_counts(preds, labs, kinds, device=None) -> dict:
    """{kind: int64 [n, 4]} for the requested kinds; kinds already in the memo are not recomputed."""
    torch = _torch()
    n = len(preds)
    if n == 0:
        return {k: np.zeros((0, 4), dtype=np.int64) for k in kinds}
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    key = (_fingerprint(preds), _fingerprint(labs), str(dev))
    if _MEMO.key != key:
        _MEMO.key, _MEMO.counts = key, {}
    todo = [k for k in kinds if k not in _MEMO.counts]
    if todo:
        h_cps, h_off, h_max = _to_device(preds, dev)
        r_cps, r_off, r_max = _to_device(labs, dev)
        modes = {"chars": MODE_CHARS, "words": MODE_WORDS, "tokens": MODE_TOKENS}
        outs = [edit_counts_device(r_cps, r_off, h_cps, h_off, n, modes[k], max(r_max, h_max), dev) for k in todo]
        # one read-back for everything that was queued
        sdih = torch.stack([o[0] for o in outs]).cpu().numpy().astype(np.int64)
        status = torch.stack([o[1] for o in outs]).cpu().numpy()
        if (status == 2).any():
            _MEMO.clear()
            i = int(np.nonzero((status == 2).any(axis=0))[0][0])
            raise _lib.CoralError(_lib.ECAP, f"pair {i}: a string is longer than the declared maximum, or the pair is "
                                             "outside rapidfuzz's direct-alignment range (len1 * len2 >= 2^22)")
        if EMPTY_REFERENCE == "raise" and (status == 1).any():
            _MEMO.clear()
            raise ValueError("one or more references are empty strings")
        for k, a in zip(todo, sdih):
            _MEMO.counts[k] = a
    return {k: _MEMO.counts[k] for k in kinds}


def edit_counts(predictions: c.Iterable[str], labels: c.Iterable[str], kind: str = "chars",
                device=None) -> np.ndarray:
    """Per-pair ``[S, D, I, H]`` (int64 ``[n, 4]``), reference = label, hypothesis = prediction.

    ``kind``: "chars" (jiwer cer_default), "words" (jiwer wer_default) or "tokens".
    Pairs are formed with ``zip`` like the reference (the shorter iterable wins).
    """
    preds, labs = _as_lists(predictions, labels)
    return _pair_counts(preds, labs, (kind,), device)[kind]


def _rate_from_counts(sdih: np.ndarray, normalise: bool) -> float:
    S, D, I, H = (int(x) for x in sdih.sum(axis=0)) if len(sdih) else (0, 0, 0, 0)
    incorrect = S + D + I
    total = S + D + H
    if normalise:
        total += I
    return incorrect / total


def cer(predictions: c.Iterable[str], labels: c.Iterable[str], normalise: bool = True) -> float:
    """Character error rate, aggregated (R:src/coral/metrics.py:8-33)."""
    preds, labs = _as_lists(predictions, labels)
    return _rate_from_counts(_pair_counts(preds, labs, ("chars", "words"))["chars"], normalise)


def wer(predictions: c.Iterable[str], labels: c.Iterable[str], normalise: bool = True) -> float:
    """Word error rate, aggregated (R:src/coral/metrics.py:36-61)."""
    preds, labs = _as_lists(predictions, labels)
    return _rate_from_counts(_pair_counts(preds, labs, ("words", "chars"))["words"], normalise)


def per_sample_rates(sdih: np.ndarray, normalise: bool = True) -> np.ndarray:
    """Each sample scored alone with the reference's formula (float64 [n]).

    This is the per-row ``asr_cer`` / ``asr_wer`` that R:src/coral/validation.py:149-158
    needs. A sample whose total is 0 cannot occur: empty references raise.
    """
    S, D, I, H = sdih[:, 0], sdih[:, 1], sdih[:, 2], sdih[:, 3]
    incorrect = S + D + I
    total = S + D + H + (I if normalise else 0)
    return incorrect / total
