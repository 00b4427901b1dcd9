"""Importable ``pyctcdecode`` namespace backed by coral_b200.

Placed on ``sys.path`` by ``coral_b200.install_shims()`` so that the unmodified HF
``Wav2Vec2ProcessorWithLM`` and ASR pipeline, which only ``import pyctcdecode`` / ``kenlm``
(HF:utils/import_utils.py:50-53; HF:pipelines/__init__.py:918-933), pick up the CUDA decoder.
"""
from coral_b200.alphabet import Alphabet  # noqa: F401
from coral_b200.decoder import BeamSearchDecoderCTC, build_ctcdecoder  # noqa: F401
from coral_b200.language_model import LanguageModel  # noqa: F401

__version__ = "0.5.0"
