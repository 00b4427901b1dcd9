"""Utterance-sharded evaluation across the GPUs of one box.

Utterances are independent from logits to per-utterance edit counts, so each rank (one
process per GPU, ``torch.distributed``) decodes and scores its own shard with no
data-path collective; the only exchange is ONE ``all_reduce(SUM)`` of an ``int64 [G, 2, 4]``
tensor -- (S, D, I, H) for characters and for words, per demographic group G -- after
which every rank divides the same integers and holds bit-identical CER / WER. A few
hundred bytes over NVLink: latency-bound, bandwidth is irrelevant.

The reference has no equivalent (its only multi-GPU use is inside accelerate/DeepSpeed
during training, R:makefile:79-137); this is the driver that splits ``evaluate`` /
``add_validations`` (R:src/coral/evaluate.py:56-84, R:src/coral/validation.py:114-140).
"""

from __future__ import annotations

import numpy as np


def shard_indices(lengths, rank: int, world_size: int) -> np.ndarray:
    """Length-balanced shard: sort by frame count (longest first), deal round-robin.
    The union over ranks is a partition of ``range(len(lengths))``."""
    lengths = np.asarray(lengths)
    order = np.argsort(-lengths, kind="stable")
    return np.sort(order[rank::world_size])


def reduce_counts(char_counts: np.ndarray, word_counts: np.ndarray, group_ids=None, n_groups: int = 1,
                  process_group=None, device=None):
    """Sum per-pair ``[n, 4]`` counts into ``int64 [n_groups, 2, 4]`` and all-reduce it.

    Works with any initialised ``torch.distributed`` backend (NCCL on the GPUs, gloo in the
    CPU tests); without an initialised process group it is the local sum.
    """
    import torch
    import torch.distributed as dist

    local = np.zeros((n_groups, 2, 4), dtype=np.int64)
    if group_ids is None:
        local[0, 0] = char_counts.sum(axis=0) if len(char_counts) else 0
        local[0, 1] = word_counts.sum(axis=0) if len(word_counts) else 0
    else:
        group_ids = np.asarray(group_ids)
        for k in range(4):
            local[:, 0, k] = np.bincount(group_ids, weights=char_counts[:, k], minlength=n_groups).astype(np.int64)
            local[:, 1, k] = np.bincount(group_ids, weights=word_counts[:, k], minlength=n_groups).astype(np.int64)
    t = torch.from_numpy(local)
    if dist.is_available() and dist.is_initialized():
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(process_group) == "nccl" \
                else torch.device("cpu")
        t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=process_group)
        t = t.cpu()
    return t.numpy()


def rates_from_totals(totals: np.ndarray, normalise: bool = True):
    """``[G, 2, 4]`` -> (cer [G], wer [G]) with the reference's formula
    (R:src/coral/metrics.py:29-33): Python-int true division, so every rank gets the same bits."""
    cers, wers = [], []
    for g in range(totals.shape[0]):
        for out, row in ((cers, totals[g, 0]), (wers, totals[g, 1])):
            S, D, I, H = (int(x) for x in row)
            total = S + D + H + (I if normalise else 0)
            out.append((S + D + I) / total if total else float("nan"))
    return cers, wers


def sharded_error_rates(predictions, labels, normalise: bool = True, process_group=None) -> dict:
    """CER / WER over the union of all ranks' (prediction, label) shards."""
    from .metrics import _as_lists, _pair_counts

    predictions, labels = _as_lists(predictions if isinstance(predictions, list) else list(predictions),
                                    labels if isinstance(labels, list) else list(labels))
    both = _pair_counts(predictions, labels, ("chars", "words"))
    cc, wc = both["chars"], both["words"]
    totals = reduce_counts(cc, wc, process_group=process_group)
    cers, wers = rates_from_totals(totals, normalise)
    return dict(cer=cers[0], wer=wers[0], totals=totals)
