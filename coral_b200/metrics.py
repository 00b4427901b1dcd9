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
# R:src/coral/evaluate.py:195-198): the code points marshalled for cer() are reused by the wer()
# that follows. cer() starts a fresh cache, so nothing survives from one evaluation to the next.
_UPLOAD_CACHE: list = []


def _upload(strings, dev):
    torch = _torch()
    # content fingerprint: str hashes are cached by CPython, so this is ~0.3 ms for 8k strings
    sig = (len(strings), hash(tuple(strings)), str(dev))
    for k, v in _UPLOAD_CACHE:
        if k == sig:
            return v[0], v[1], v[2]
    cps, off = encode_utf32(strings)
    max_len = int(np.diff(off).max()) if len(off) > 1 else 0
    d_cps = torch.from_numpy(cps.view(np.int32)).to(dev, non_blocking=True)
    d_off = torch.from_numpy(off).to(dev, non_blocking=True)
    _UPLOAD_CACHE.append((sig, (d_cps, d_off, max_len)))
    del _UPLOAD_CACHE[:-4]
    return d_cps, d_off, max_len


def edit_counts(predictions: c.Iterable[str], labels: c.Iterable[str], kind: str = "chars",
                device=None) -> np.ndarray:
    """Per-pair ``[S, D, I, H]`` (int64 ``[n, 4]``), reference = label, hypothesis = prediction.

    ``kind``: "chars" (jiwer cer_default), "words" (jiwer wer_default) or "tokens".
    Pairs are formed with ``zip`` like the reference (the shorter iterable wins).
    """
    torch = _torch()
    if isinstance(predictions, list) and isinstance(labels, list) and len(predictions) == len(labels):
        preds, labs = predictions, labels
    else:
        pairs = list(zip(predictions, labels))
        preds = [p for p, _ in pairs]
        labs = [l for _, l in pairs]
    n = len(preds)
    if n == 0:
        return np.zeros((0, 4), dtype=np.int64)
    for s in labs:
        if not isinstance(s, str):
            raise TypeError("references must be strings")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    mode = {"chars": MODE_CHARS, "words": MODE_WORDS, "tokens": MODE_TOKENS}[kind]
    r_cps, r_off, r_max = _upload(labs, dev)
    h_cps, h_off, h_max = _upload(preds, dev)
    sdih, status = edit_counts_device(r_cps, r_off, h_cps, h_off, n, mode, max(r_max, h_max), dev)
    sdih = sdih.cpu().numpy().astype(np.int64)
    if status.cpu().numpy().any():
        raise ValueError("one or more references are empty strings")
    return sdih


def _rate_from_counts(sdih: np.ndarray, normalise: bool) -> float:
    S, D, I, H = (int(x) for x in sdih.sum(axis=0)) if len(sdih) else (0, 0, 0, 0)
    incorrect = S + D + I
    total = S + D + H
    if normalise:
        total += I
    return incorrect / total


def cer(predictions: c.Iterable[str], labels: c.Iterable[str], normalise: bool = True) -> float:
    """Character error rate, aggregated (R:src/coral/metrics.py:8-33)."""
    _UPLOAD_CACHE.clear()
    return _rate_from_counts(edit_counts(predictions, labels, "chars"), normalise)


def wer(predictions: c.Iterable[str], labels: c.Iterable[str], normalise: bool = True) -> float:
    """Word error rate, aggregated (R:src/coral/metrics.py:36-61)."""
    return _rate_from_counts(edit_counts(predictions, labels, "words"), normalise)


def per_sample_rates(sdih: np.ndarray, normalise: bool = True) -> np.ndarray:
    """Each sample scored alone with the reference's formula (float64 [n]).

    This is the per-row ``asr_cer`` / ``asr_wer`` that R:src/coral/validation.py:149-158
    needs. A sample whose total is 0 cannot occur: empty references raise.
    """
    S, D, I, H = sdih[:, 0], sdih[:, 1], sdih[:, 2], sdih[:, 3]
    incorrect = S + D + I
    total = S + D + H + (I if normalise else 0)
    return incorrect / total
