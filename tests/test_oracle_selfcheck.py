"""Pins the oracle to the real pyctcdecode / kenlm / jiwer wherever they are importable
(they are not in the build image: the test then skips -- SURVEY.md section 8c "upgrade path")."""

import numpy as np
import pytest

from oracle import selfcheck


@pytest.mark.skipif(not selfcheck.real_packages_available(), reason="pyctcdecode/kenlm/jiwer are not installed here")
def test_oracle_matches_the_real_packages(small_lm, small_workload):
    import synth

    rng = np.random.default_rng(5)
    pairs = [("ab", "ba"), ("abc", "bcd"), ("hej med dig", "hej  med   dig"), ("a b c d", "a x c")]
    for _ in range(300):
        r = "".join(rng.choice(list("abcde "), size=int(rng.integers(1, 40)))).strip() or "a"
        h = "".join(rng.choice(list("abcde "), size=int(rng.integers(0, 40))))
        pairs.append((r, h))
    selfcheck.check_edit(pairs)
    w = small_workload
    selfcheck.check_beam(synth.CORAL_LABELS, small_lm[2], [w.logits[u, : w.lengths[u]] for u in range(4)])
