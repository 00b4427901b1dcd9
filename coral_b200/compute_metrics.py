"""Drop-in for ``coral.compute_metrics.compute_error_rate_metrics``
(R:src/coral/compute_metrics.py:18-94), the hook ``Trainer.evaluate`` calls
(R:src/coral/wav2vec2.py:152-154, :304).

Same argument meaning, same in-place ``labels[labels == -100] = pad`` fix-up (:46-48),
same "-100 row -> pad" rule (:66), first-maximum argmax (:68), grouped decode of the
predictions and ungrouped decode of the labels (:69-70), ``.lower().strip()`` (:80-81)
and the ``{"cer", "wer"}`` result (:90-94). The argmax/collapse and the edit counts run
on the GPU; the Whisper (2-D ids) branch keeps the reference's host decoding and only
its CER/WER are accelerated.
"""

from __future__ import annotations

import logging
import os

import numpy as np

from .greedy import CTCVocabulary, decode_ids, greedy_decode
from .metrics import cer, wer

logger = logging.getLogger(__name__)


def compute_error_rate_metrics(pred, processor, log_examples: bool = True) -> dict[str, float]:
    tokenizer = getattr(processor, "tokenizer")
    pad_token = tokenizer.pad_token_id
    predictions = pred.predictions
    labels = pred.label_ids
    assert isinstance(labels, np.ndarray)
    labels[labels == -100] = pad_token

    if predictions.ndim == 2:
        # Whisper ids: host decoding exactly as the reference does (:52-59)
        if type(processor).__name__ == "Wav2Vec2ProcessorWithLM":
            predictions_str = processor.batch_decode(predictions)
        else:
            predictions_str = processor.batch_decode(predictions, skip_special_tokens=True)
        labels_str = tokenizer.batch_decode(sequences=labels, skip_special_tokens=True)
    elif predictions.ndim == 3:
        vocab = CTCVocabulary.from_tokenizer(tokenizer)
        predictions_str = greedy_decode(predictions, vocab, lengths=None, pad_fixup=True)
        labels_str = decode_ids(labels, vocab, group_tokens=False)
    else:
        raise ValueError(
            f"Expected predictions to have either 2 or 3 dimensions, but found "
            f"{predictions.ndim} dimensions."
        )

    predictions_str = [p.lower().strip() for p in predictions_str]
    labels_str = [lbl.lower().strip() for lbl in labels_str]

    is_main_process = os.getenv("RANK", "0") == "0"
    if is_main_process and log_examples:
        random_idx = np.random.randint(0, len(predictions_str))
        logger.info(f"Random sample document: {labels_str[random_idx]}")
        logger.info(f"Associated prediction: {predictions_str[random_idx]}")

    return dict(
        cer=cer(predictions=predictions_str, labels=labels_str),
        wer=wer(predictions=predictions_str, labels=labels_str),
    )
