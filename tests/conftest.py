"""Shared fixtures. ``-m "not gpu"`` runs on CPU (oracle, host logic, ABI exports, host
simulation of the kernel logic); ``-m gpu`` tests are the parity tests proper and call the
CUDA library through the C ABI. Nothing here reads /root/reference at run time."""

from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

CACHE = os.environ.get("CORAL_B200_CACHE", os.path.join(tempfile.gettempdir(), "coral_b200_cache"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cache_dir():
    os.makedirs(CACHE, exist_ok=True)
    return CACHE


@pytest.fixture(scope="session")
def small_lm(cache_dir):
    """(words, corpus model, ARPA path) of a small 4-gram LM (2k words, 5k sentences)."""
    import synth

    return synth.build_lm(cache_dir, order=4, n_words=2000, n_sent=5000)


@pytest.fixture(scope="session")
def small_workload(cache_dir, small_lm):
    import synth

    return synth.build_workload(cache_dir, 12, order=4, n_words=2000, n_sent=5000, name="t")


@pytest.fixture(scope="session")
def oracle_decoder(small_lm):
    import synth
    from oracle.beam import build_ctcdecoder

    return build_ctcdecoder(synth.CORAL_LABELS, small_lm[2])


# how often the near-tie licence of beams_equal fired in this session (printed at the end)
TIE_STATS = {"calls": 0, "beams": 0, "swaps": 0, "max_gap": 0.0}


def beams_equal(ref_beams, got_beams, rel=1e-4, tie=1e-5):
    """ref: oracle 5-tuples; got: (text, ..., logit, lm) with scores at [-2], [-1].

    Texts, word frames and the ORDER of the beams must agree, scores within ``rel``. One licence:
    CUDA's expf/logf and numpy's differ in the last ulp of the float32 log-softmax, so two beams
    whose combined scores the reference itself separates by less than ``tie`` (absolute, about ten
    times the accumulated ulp noise of a short utterance) may come out in the other order; they
    are matched by text and everything else is still compared. The licence never applies to the
    best beam (rank 0 is the transcript the callers use: it must be identical, full stop).
    Returns the number of beams matched out of position; the session total is reported at the end
    of the run (``TIE_STATS``)."""
    assert len(ref_beams) == len(got_beams), (len(ref_beams), len(got_beams))
    where = {g[0]: k for k, g in enumerate(got_beams)}
    assert len(where) == len(got_beams), "duplicate transcripts in the beam list"
    swaps = 0
    for i, (r, g) in enumerate(zip(ref_beams, got_beams)):
        if r[0] != g[0]:
            assert i > 0, f"best beam differs: {r[0]!r} != {g[0]!r}"
            assert r[0] in where, f"beam {i}: {r[0]!r} missing (got {g[0]!r})"
            j = where[r[0]]
            assert j > 0, f"beam {i}: {r[0]!r} was returned as the best beam"
            gap = abs(ref_beams[i][-1] - ref_beams[j][-1])
            assert gap <= tie, f"beam {i}: {r[0]!r} != {g[0]!r} and no near-tie (gap {gap})"
            g = got_beams[j]
            swaps += 1
            TIE_STATS["max_gap"] = max(TIE_STATS["max_gap"], gap)
        assert abs(r[-2] - g[-2]) <= rel * max(1.0, abs(r[-2])), (i, r[-2], g[-2])
        assert abs(r[-1] - g[-1]) <= rel * max(1.0, abs(r[-1])), (i, r[-1], g[-1])
        if len(g) >= 4:  # word frames travel with the beam: (word, (start_frame, end_frame))
            rf = [(w, (int(a), int(b))) for w, (a, b) in r[2]]
            gf = [(w, (int(a), int(b))) for w, (a, b) in g[-3]]
            assert rf == gf, f"beam {i} ({r[0]!r}): word frames {rf} != {gf}"
    TIE_STATS["calls"] += 1
    TIE_STATS["beams"] += len(ref_beams)
    TIE_STATS["swaps"] += swaps
    return swaps


def pytest_terminal_summary(terminalreporter):
    if TIE_STATS["calls"]:
        terminalreporter.write_line(
            "beams_equal: %d beam lists, %d beams compared, %d matched out of position under the near-tie "
            "licence (never the best beam; largest oracle gap %.2e)"
            % (TIE_STATS["calls"], TIE_STATS["beams"], TIE_STATS["swaps"], TIE_STATS["max_gap"]))


@pytest.fixture
def rng(request):
    """A generator per test, seeded by the test's id: the inputs of a test do not depend on which
    other tests ran before it."""
    import zlib

    return np.random.default_rng([20261017, zlib.crc32(request.node.nodeid.encode())])
