"""Cross-check of the oracle against the REAL packages, on any machine where they import.

pyctcdecode / kenlm / jiwer are not installable in the build image (oracle/__init__.py), so the
restatements are "parity unpinned". Wherever the real packages are importable this module pins
them: same transcripts and scores for beam search, same (S, D, I, H) for jiwer. It is run by
tests/test_oracle_selfcheck.py, which skips when the imports fail.
"""

from __future__ import annotations

import numpy as np


def real_packages_available() -> bool:
    import importlib.util
    import sys

    # the coral_b200 shims are not the real thing
    for name in ("pyctcdecode", "kenlm", "jiwer"):
        mod = sys.modules.get(name)
        spec = importlib.util.find_spec(name) if mod is None else getattr(mod, "__spec__", None)
        if spec is None or "coral_b200" in str(getattr(spec, "origin", "")):
            return False
    return True


def check_edit(pairs) -> None:
    import jiwer

    from . import edit

    for ref, hyp in pairs:
        m = jiwer.process_characters(reference=ref, hypothesis=hyp)
        assert edit.char_counts(ref, hyp) == (m.substitutions, m.deletions, m.insertions, m.hits), (ref, hyp)
        m = jiwer.process_words(reference=ref, hypothesis=hyp)
        assert edit.word_counts(ref, hyp) == (m.substitutions, m.deletions, m.insertions, m.hits), (ref, hyp)


def check_beam(labels, arpa_path, logits_list, **kwargs) -> None:
    from pyctcdecode import build_ctcdecoder as real_build

    from .beam import build_ctcdecoder

    real = real_build(list(labels), kenlm_model_path=arpa_path)
    ours = build_ctcdecoder(list(labels), arpa_path)
    for lg in logits_list:
        a = real.decode_beams(np.asarray(lg), **kwargs)
        b = ours.decode_beams(np.asarray(lg), **kwargs)
        assert [x[0] for x in a] == [x[0] for x in b]
        for x, y in zip(a, b):
            assert abs(x[3] - y[3]) <= 1e-6 * max(1.0, abs(y[3])) and abs(x[4] - y[4]) <= 1e-6 * max(1.0, abs(y[4]))
