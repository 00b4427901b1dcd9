"""CPU check of the beam-search KERNEL LOGIC (coral_b200/csrc/beam_core.h compiled with
-DCORAL_HOSTSIM, lanes run sequentially) against the oracle. This validates the device
algorithm -- prefix identity, gather-merge, LM records, prune/trim/rank, overflow path --
without a GPU; the `-m gpu` tests then check the real kernel through the C ABI."""

from __future__ import annotations

import numpy as np
import pytest

from conftest import beams_equal


@pytest.fixture(scope="module")
def sim(small_lm, oracle_decoder):
    from hostsim_lib import HostSim

    return HostSim(oracle_decoder._alphabet.labels, small_lm[2])


def _check(dec, hs, lg, **kw):
    okw = {k: v for k, v in kw.items() if k in ("beam_width", "beam_prune_logp", "token_min_logp")}
    dec.reset_params(alpha=kw.get("alpha", 0.5), beta=kw.get("beta", 1.5),
                     unk_score_offset=kw.get("unk_score_offset", -10.0),
                     lm_score_boundary=kw.get("score_boundary", True))
    try:
        ref = dec.decode_beams(lg, **okw)
    finally:
        dec.reset_params(alpha=0.5, beta=1.5, unk_score_offset=-10.0, lm_score_boundary=True)
    got = hs.decode_beams(lg, **kw)
    beams_equal(ref, got)


def test_lm_sentence_scores_bit_exact(sim, oracle_decoder, small_lm, rng):
    import synth

    words, model, _ = small_lm
    m = oracle_decoder._language_model._kenlm_model
    flat, lens = model.sample(150, "hs-lm")
    for s in synth.sentences_to_text(flat, lens, words):
        ws = s.split(" ")
        if rng.random() < 0.5:
            ws[int(rng.integers(len(ws)))] = "zzzqq"
        st = m.begin_sentence_state()
        ref = []
        for w in ws:
            p, st = m.base_score(st, w)
            ref.append(p)
        ref.append(m.base_score(st, "</s>")[0])
        got, oov = sim.score_sentence(ws)
        assert np.array_equal(np.array(ref, dtype=np.float32), got)
        assert oov.tolist() == [int(w not in m) for w in ws]


@pytest.mark.parametrize("variant", [0, 3, 4])
def test_peaky_utterances(sim, oracle_decoder, small_workload, variant):
    w = small_workload
    for u in range(6):
        _check(oracle_decoder, sim, w.logits[u, : w.lengths[u]], variant=variant, repeat=2)


def test_parameter_sweep(sim, oracle_decoder, small_workload):
    w = small_workload
    for u in range(3):
        lg = w.logits[u, : w.lengths[u]]
        _check(oracle_decoder, sim, lg, beam_width=20, variant=1)
        _check(oracle_decoder, sim, lg, token_min_logp=-10.0, beam_prune_logp=-5.0)
        _check(oracle_decoder, sim, lg, token_min_logp=-20.0, beam_width=64)
        _check(oracle_decoder, sim, lg, alpha=0.9, beta=0.3, unk_score_offset=-4.0, score_boundary=False)


def test_flat_logits_overflow_and_trim(sim, oracle_decoder, rng):
    import synth

    flat = synth.flat_logits(50, rng)
    _check(oracle_decoder, sim, flat)                              # overflow path (4k candidates / frame)
    _check(oracle_decoder, sim, flat, beam_width=16, variant=1)
    _check(oracle_decoder, sim, flat, variant=3)                   # small shared-memory candidate arrays
    _check(oracle_decoder, sim, flat, variant=4)                   # the production default <128, 104, 208>
    _check(oracle_decoder, sim, flat[:24], beam_width=512, variant=2)
    _check(oracle_decoder, sim, flat[:30], beam_width=300, token_min_logp=-3.0, variant=2)


def test_bunched_and_tied_scores_take_the_radix_select(sim, oracle_decoder, rng):
    """Scores bunched into one ranking bucket (near-constant logits) overflow the shared
    candidate arrays even after the bucket cut, so the 64-bit radix select runs; exactly
    constant logits add bit-for-bit ties at the cut (earlier candidate wins)."""
    near = (rng.standard_normal((14, 46)) * 1e-3).astype(np.float32)
    _check(oracle_decoder, sim, near)
    assert sim.last_stats[7] > 0
    _check(oracle_decoder, sim, np.zeros((8, 46), np.float32))
    assert sim.last_stats[7] > 0
    _check(oracle_decoder, sim, np.zeros((6, 46), np.float32), beam_width=16, variant=1)


def test_word_frames_match_the_oracle(sim, oracle_decoder, small_workload, rng):
    """text_frames of every final beam (pyctcdecode's word time offsets): peaky utterances,
    flat logits (merges pick the later member's frames), trailing open word, leading space."""
    import synth

    w = small_workload
    cases = [w.logits[u, : w.lengths[u]] for u in range(5)] + [synth.flat_logits(40, rng), w.logits[0, :7],
                                                                w.logits[1, :1], np.zeros((0, 46), np.float32)]
    for lg in cases:
        ref = oracle_decoder.decode_beams(lg)
        for variant in (0, 3):
            got = sim.decode_beams(lg, frames=True, variant=variant)
            assert len(ref) == len(got)
            for r, g in zip(ref, got):
                assert r[0] == g[0]
                assert [(wd, (int(a), int(b))) for wd, (a, b) in r[2]] == g[1], (r[0], r[2], g[1])
    # the frames instantiation returns the same beams and scores as the plain one
    lg = cases[0]
    a = sim.decode_beams(lg)
    b = sim.decode_beams(lg, frames=True)
    assert [(x[0], x[1], x[2]) for x in a] == [(x[0], x[2], x[3]) for x in b]


def test_long_flat_utterance_many_prefixes(sim, oracle_decoder, rng):
    """Thousands of distinct prefixes in one utterance (back-pointer arena, slot reuse)."""
    import synth

    lg = synth.flat_logits(120, rng)
    _check(oracle_decoder, sim, lg, beam_width=100, repeat=2)
    assert sim.last_stats[5] > 2048


def test_edge_cases(sim, oracle_decoder, small_workload):
    w = small_workload
    _check(oracle_decoder, sim, np.zeros((0, 46), np.float32))
    _check(oracle_decoder, sim, w.logits[0, :1])
    _check(oracle_decoder, sim, w.logits[0, :2], beam_width=1, variant=1)
    # -100-padded tail frames decoded as frames (what pyctcdecode does if they are not stripped)
    lg = np.concatenate([w.logits[1, :40], np.full((5, 46), -100.0, np.float32)])
    _check(oracle_decoder, sim, lg)


def test_no_lm_and_unigram_variants(small_lm, small_workload, rng):
    import synth
    from hostsim_lib import HostSim
    from oracle.arpa import ArpaModel
    from oracle.beam import Alphabet, BeamSearchDecoderCTC, build_ctcdecoder
    from oracle.lm import LanguageModel

    w = small_workload
    words, _, path = small_lm
    d0 = build_ctcdecoder(synth.CORAL_LABELS)
    h0 = HostSim(d0._alphabet.labels)
    _check(d0, h0, synth.flat_logits(40, rng))
    _check(d0, h0, w.logits[0, : w.lengths[0]])
    alpha = Alphabet.build_alphabet(synth.CORAL_LABELS)
    dn = BeamSearchDecoderCTC(alpha, LanguageModel(ArpaModel.load(path), None))
    _check(dn, HostSim(dn._alphabet.labels, path, unigrams=None), w.logits[1, : w.lengths[1]])
    sub = sorted(set(words[:500]))
    ds = BeamSearchDecoderCTC(alpha, LanguageModel(ArpaModel.load(path), sub))
    _check(ds, HostSim(ds._alphabet.labels, path, unigrams=sub), w.logits[2, : w.lengths[2]])


def test_probability_inputs(sim, oracle_decoder, small_workload):
    import math

    w = small_workload
    z = w.logits[3, : w.lengths[3]].astype(np.float64)
    p = np.exp(z - z.max(axis=1, keepdims=True))
    p = (p / p.sum(axis=1, keepdims=True)).astype(np.float32)
    if not math.isclose(float(p.sum(axis=1).mean()), 1):
        pytest.skip("float32 mean of the row sums is not exactly 1 for this draw")
    _check(oracle_decoder, sim, p)
    # the small instantiations share their staging scratch with fewer candidate slots
    _check(oracle_decoder, sim, p, beam_width=16, variant=1)
    _check(oracle_decoder, sim, p, variant=3)


def test_work_counters_match_oracle(sim, oracle_decoder, small_workload):
    """Device-side work counters agree with the oracle's: beam extensions and
    frames exactly. LM word scorings are >= the oracle's cached_lm_scores misses: the device
    re-scores a completed word whose text was scored before but is no longer on any beam
    (same value, no text-keyed cache to consult)."""
    w = small_workload
    lg = w.logits[4, : w.lengths[4]]
    s0 = dict(oracle_decoder.stats)
    oracle_decoder.decode_beams(lg)
    d = {k: oracle_decoder.stats[k] - s0[k] for k in s0}
    sim.decode_beams(lg)
    st = sim.last_stats
    assert int(st[0]) == d["extensions"] and int(st[3]) == d["frames"]
    assert d["n_score"] <= int(st[1]) <= 2 * d["n_score"]


def test_pairwise_sum_is_numpys_float32_order(rng):
    """The kernel's probabilities-vs-logits test replays np.add.reduce's float32 summation
    order (pyctcdecode: logits.sum(axis=1).mean()); checked bit for bit across the leaf,
    remainder and recursive-split regimes."""
    import ctypes as C

    from hostsim_lib import build

    lib = C.CDLL(build())
    lib.hs_pairwise_sum.restype = C.c_float
    lib.hs_pairwise_sum.argtypes = [C.c_void_p, C.c_int]
    for n in list(range(0, 140)) + [255, 256, 257, 499, 500, 1000, 1023, 1500, 4097]:
        for scale in (1.0, 1e-3):
            a = (rng.standard_normal(n) * 50 * scale + scale).astype(np.float32)
            want = np.add.reduce(a) if n else np.float32(0)
            got = lib.hs_pairwise_sum(a.ctypes.data, n)
            assert np.float32(got) == np.float32(want), (n, got, want)


# ---- the named BASELINE configs through the production instantiations (tests/cells.py) --------
def _variant_for(beam_width):
    # coral_b200/csrc/beam.cu: text-only instantiation per beam width
    return 1 if beam_width <= 32 else 8 if beam_width <= 64 else 4 if beam_width <= 104 else \
        7 if beam_width <= 128 else 5 if beam_width <= 256 else 6


@pytest.mark.parametrize("cell,utts", [
    ("c2flat", (10, 15)), ("c3_tml20", (15,)), ("c3flat_tml3", (14, 15)), ("c3flat_tml5", (15,)),
    ("c5_o3_b16", (0, 8)), ("c5_o4_b64", (0, 8)), ("c5_o5_b128", (0, 8)), ("c5_o6_b256", (0, 8)),
    ("c5_o5_b512", (0, 8)), ("c3_tml5", (0,)), ("c3_tml10", (15,))])
def test_config_cells_device_logic(cell, utts):
    """Kernel logic (host simulation) == oracle on utterances of the config cells, all beams, text
    and word frames; also pins the committed cell goldens to the live inputs."""
    import cells
    from hostsim_lib import HostSim
    from oracle.beam import Alphabet

    labels, arpa, logits, kw = cells.cell_inputs(cell)
    if cell in cells.CACHED:
        data = cells.load_cache(cell)
        assert data is not None and data["fingerprint"] == cells.fingerprint(cell), \
            f"tests/golden/cells/{cell}.json.gz is stale: re-run tests/golden/make_cells.py"
        ref_all = [[(t, None, [(w, (a, b)) for w, (a, b) in fr], ls, cs) for t, fr, ls, cs in utt]
                   for utt in data["beams"]]
        ref = [ref_all[u] for u in utts]
    else:
        live = cells.live_oracle(cell, which=utts, n_procs=1)
        ref = [[(t, None, [(w, (a, b)) for w, (a, b) in fr], ls, cs) for t, fr, ls, cs in utt] for utt in live]
    key = ("cells", arpa)
    if key not in _SIMS:
        _SIMS[key] = HostSim(Alphabet.build_alphabet(list(labels)).labels, arpa)
    hs = _SIMS[key]
    for u, r in zip(utts, ref):
        for frames in (False, True):
            got = hs.decode_beams(logits[u], variant=_variant_for(kw["beam_width"]), frames=frames, **kw)
            beams_equal(r, got)


_SIMS = {}


@pytest.mark.parametrize("variant", [4, 0, 8])
def test_prune_history_matches_the_oracle(sim, oracle_decoder, small_workload, rng, variant):
    """pyctcdecode prune_history=True (first beam per (last order-1 words, word_part, last_char) after
    every trim) in the kernel logic == the oracle's _prune_history; with and without an LM."""
    import synth
    from hostsim_lib import HostSim

    w = small_workload
    bw = 64 if variant == 8 else 100
    for u in range(5):
        lg = w.logits[u, : w.lengths[u]]
        ref = oracle_decoder.decode_beams(lg, beam_width=bw, prune_history=True)
        for frames in (False, True):
            beams_equal(ref, sim.decode_beams(lg, beam_width=bw, variant=variant, prune_history=True, frames=frames))
        assert len(ref) <= len(oracle_decoder.decode_beams(lg, beam_width=bw))
    flat = synth.flat_logits(40, rng)
    beams_equal(oracle_decoder.decode_beams(flat, beam_width=bw, prune_history=True),
                sim.decode_beams(flat, beam_width=bw, variant=variant, prune_history=True))
    if variant == 4:
        from oracle.beam import build_ctcdecoder as oracle_build

        o = oracle_build(synth.CORAL_LABELS)
        hs = HostSim(o._alphabet.labels)
        for lg in (w.logits[0, : w.lengths[0]], flat):
            beams_equal(o.decode_beams(lg, prune_history=True), hs.decode_beams(lg, variant=4, prune_history=True))
