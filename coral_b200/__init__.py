"""coral_b200 -- B200-native CTC decoding and WER/CER scoring behind CoRal's Python surface.

Hot path only: greedy argmax-collapse, pyctcdecode-style prefix beam search
with KenLM n-gram shallow fusion, and Levenshtein WER/CER, as hand-written CUDA for sm_100a
behind the C ABI in ``include/coral_b200.h``. There is no CPU fallback: importing the
sub-modules is cheap, but every compute entry point needs the built library and a GPU.
"""
from __future__ import annotations

import os
import sys

__version__ = "0.1.0"

SHIMS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def install_shims() -> None:
    """Make ``import pyctcdecode`` / ``import kenlm`` resolve to the coral_b200 shims."""
    if SHIMS_DIR not in sys.path:
        sys.path.insert(0, SHIMS_DIR)
    for name in ("pyctcdecode", "kenlm"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith(SHIMS_DIR):
            for k in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
                del sys.modules[k]
