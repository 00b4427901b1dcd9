"""Greedy CTC decode and the Trainer metric hook, restated in numpy (test infrastructure only).

* ``ids_to_string`` restates ``Wav2Vec2CTCTokenizer._decode`` /
  ``convert_tokens_to_string`` (HF:models/wav2vec2/tokenization_wav2vec2.py:410-459,
  :296-357) with ``skip_special_tokens=False``: id -> token (unknown id ->
  ``unk_token``), ``itertools.groupby`` collapse when ``group_tokens``, drop
  tokens equal to ``pad_token``, word delimiter -> ``" "``, join, ``strip()``.
  This one IS pinned: tests compare it with the real tokenizer of the
  transformers version the reference pins (R:uv.lock:3290-3291).
* ``compute_error_rate_metrics`` restates R:src/coral/compute_metrics.py:18-94 for
  the wav2vec2 (3-D) branch.
"""

from __future__ import annotations

import itertools

import numpy as np

from . import edit

# the CoRal alphabet: R:src/coral/wav2vec2.py:318-322 with
# characters_to_keep from R:config/model/wav2vec2-small.yaml:10, then the four
# special tokens the tokenizer appends (R:src/coral/wav2vec2.py:64-72).
CORAL_CHARACTERS = "abcdefghijklmnopqrstuvwxyzæøå0123456789éü"
CORAL_VOCAB = sorted(set(CORAL_CHARACTERS + "|")) + ["<s>", "</s>", "<unk>", "<pad>"]
CORAL_PAD_ID = 45
CORAL_DELIM_ID = 36


def ids_to_string(ids, vocab=CORAL_VOCAB, pad_token="<pad>", word_delimiter="|",
                  unk_token="<unk>", group_tokens: bool = True) -> str:
    tokens = [vocab[i] if 0 <= i < len(vocab) else unk_token for i in ids]
    if group_tokens:
        tokens = [t for t, _ in itertools.groupby(tokens)]
    tokens = [t for t in tokens if t != pad_token]
    return "".join(" " if t == word_delimiter else t for t in tokens).strip()


def greedy_ids(logits: np.ndarray) -> np.ndarray:
    """``np.argmax(axis=-1)``: first maximum wins (R:src/coral/compute_metrics.py:68)."""
    return np.argmax(logits, axis=-1)


def compute_error_rate_metrics(predictions: np.ndarray, label_ids: np.ndarray,
                               vocab=CORAL_VOCAB, pad_id: int = CORAL_PAD_ID) -> dict[str, float]:
    predictions = np.array(predictions, copy=True)
    labels = np.array(label_ids, copy=True)
    labels[labels == -100] = pad_id
    if predictions.ndim != 3:
        raise ValueError(
            f"Expected predictions to have either 2 or 3 dimensions, but found "
            f"{predictions.ndim} dimensions."
        )
    predictions[np.all(predictions == -100, axis=-1), pad_id] = 0
    pred_ids = np.argmax(predictions, axis=-1)
    predictions_str = [ids_to_string(row, vocab) for row in pred_ids]
    labels_str = [ids_to_string(row, vocab, group_tokens=False) for row in labels]
    predictions_str = [p.lower().strip() for p in predictions_str]
    labels_str = [l.lower().strip() for l in labels_str]
    return dict(cer=edit.cer(predictions_str, labels_str), wer=edit.wer(predictions_str, labels_str))
