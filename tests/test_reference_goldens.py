"""Consumes the REFERENCE-generated goldens ``tests/golden/ref_*.json`` that tools/pin_oracle.py
writes on a machine where the real pyctcdecode / kenlm / jiwer import. They do not exist in
this image (the packages are not installable: oracle/__init__.py), so these tests skip here;
the day one machine has the packages and the files are committed, the oracle (CPU) and the CUDA
path (GPU) are pinned to the real packages' outputs on every run."""

from __future__ import annotations

import json
import os

import numpy as np
import pytest

from conftest import beams_equal

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _ref(name):
    p = os.path.join(G, name)
    if not os.path.exists(p):
        pytest.skip(f"{name} not generated yet (tools/pin_oracle.py needs the real packages)")
    with open(p, encoding="utf-8") as f:
        return json.load(f)


def _beam_inputs(data, small_workload):
    import hashlib

    import synth

    w = small_workload
    rng = np.random.default_rng(5)
    flats = [synth.flat_logits(40 + 20 * k, rng) for k in range(3)]
    out = []
    for c in data["cases"]:
        n = c["name"]
        lg = flats[int(n[4:])] if n.startswith("flat") else w.logits[int(n[1:].split("_")[0])]
        if not n.startswith("flat"):
            lg = lg[: w.lengths[int(n[1:].split("_")[0])]]
        assert hashlib.sha256(np.ascontiguousarray(lg).tobytes()).hexdigest() == c["sha256"], "inputs drifted: " + n
        ref = [(t, None, [(wd, (a, b)) for wd, (a, b) in fr], ls, cs) for t, fr, ls, cs in c["beams"]]
        out.append((n, lg, c["kwargs"], ref))
    return out


def test_oracle_matches_reference_generated_goldens(oracle_decoder, small_workload):
    from oracle import edit as oe

    ed = _ref("ref_edit.json")
    for r in ed["rows"]:
        assert list(oe.char_counts(r["ref"], r["hyp"])) == r["chars"], r
        assert list(oe.word_counts(r["ref"], r["hyp"])) == r["words"], r
    for n, lg, kw, ref in _beam_inputs(_ref("ref_beam.json"), small_workload):
        beams_equal(ref, oracle_decoder.decode_beams(lg, **kw), rel=1e-6, tie=0.0)
    m = oracle_decoder._language_model._kenlm_model
    for row in _ref("ref_lm.json")["rows"]:
        st = m.begin_sentence_state()
        got = []
        for wd in row["sentence"].split(" "):
            x, st = m.base_score(st, wd)
            got.append(float(np.float32(x)))
        got.append(float(np.float32(m.base_score(st, "</s>")[0])))
        assert got == [float(np.float32(p)) for p, _, _ in row["full_scores"]], row["sentence"]


@pytest.mark.gpu
def test_cuda_path_matches_reference_generated_goldens(small_lm, small_workload):
    import synth
    from coral_b200.decoder import build_ctcdecoder
    from coral_b200.metrics import edit_counts

    ed = _ref("ref_edit.json")
    rows = [r for r in ed["rows"]]
    cc = edit_counts([r["hyp"] for r in rows], [r["ref"] for r in rows], "chars")
    wc = edit_counts([r["hyp"] for r in rows], [r["ref"] for r in rows], "words")
    assert cc.tolist() == [r["chars"] for r in rows] and wc.tolist() == [r["words"] for r in rows]
    dec = build_ctcdecoder(synth.CORAL_LABELS, small_lm[2])
    for n, lg, kw, ref in _beam_inputs(_ref("ref_beam.json"), small_workload):
        beams_equal(ref, dec.decode_beams_batch(None, [lg], **kw)[0])
