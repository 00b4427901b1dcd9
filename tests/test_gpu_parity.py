"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on
the same seeded inputs. Integer / byte / index results are bit-exact; beam-search
transcripts identical with scores within 1e-4 relative (the tolerance BASELINE.json states)."""

from __future__ import annotations

import os

import numpy as np
import pytest

from conftest import beams_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.fixture(scope="module")
def gpu_decoder(small_lm, torch_cuda):
    import synth
    from coral_b200.decoder import build_ctcdecoder

    return build_ctcdecoder(synth.CORAL_LABELS, small_lm[2])


@pytest.fixture(scope="module")
def coral_vocab():
    from coral_b200.greedy import CTCVocabulary
    from oracle.greedy import CORAL_PAD_ID, CORAL_VOCAB

    return CTCVocabulary(CORAL_VOCAB, CORAL_PAD_ID)


# ------------------------------------------------------------------------------ greedy
def test_greedy_matches_numpy_argmax_and_tokenizer_rules(torch_cuda, coral_vocab, rng):
    from coral_b200.greedy import greedy_decode_device
    from oracle.greedy import ids_to_string

    torch = torch_cuda
    B, T, V = 64, 499, 46  # config 1 of BASELINE.json
    logits = (rng.standard_normal((B, T, V)) * 2).astype(np.float32)
    # exact ties (first maximum must win), special tokens winning, and repeated frames
    logits[0, :50, 3] = logits[0, :50, 7] = 9.0
    logits[1, :40, 42] = 12.0
    logits[2, 10:60] = logits[2, 10]
    lengths = rng.integers(1, T + 1, size=B).astype(np.int32)
    lengths[3] = T
    d = torch.from_numpy(logits).cuda()
    ids, tokens, lens = greedy_decode_device(d, torch.from_numpy(lengths).cuda(), blank_id=45, want_ids=True)
    ids, tokens, lens = ids.cpu().numpy(), tokens.cpu().numpy(), lens.cpu().numpy()
    ref_ids = np.argmax(logits, axis=-1)
    got = coral_vocab.to_strings(tokens, lens)
    for b in range(B):
        assert np.array_equal(ids[b, : lengths[b]], ref_ids[b, : lengths[b]])
        assert got[b] == ids_to_string(ref_ids[b, : lengths[b]])


@pytest.mark.parametrize("V,T,offset", [(46, 300, 0), (46, 257, 2), (45, 129, 0), (7, 40, 1), (64, 128, 0)])
def test_greedy_argmax_special_values_and_alignments(torch_cuda, rng, V, T, offset):
    """np.argmax semantics on the awkward values (first NaN wins, -inf rows, +-0 ties, ties
    between even and odd positions) for even and odd vocabularies, and for a logits buffer that
    starts 4 * offset bytes off 16-byte alignment (the bulk-copy path must stay inside it)."""
    from coral_b200 import _lib

    torch = torch_cuda
    B = 9
    x = (rng.standard_normal((B, T, V)) * 3).astype(np.float32)
    x[0, :, :] = -np.inf
    x[1, ::3, 5 % V] = np.nan
    x[1, ::6, 2 % V] = np.nan
    x[2, :, :] = 0.0
    x[2, ::2, 3 % V] = -0.0
    x[3, :, 1] = x[3, :, 4 % V] = 50.0           # odd position ties with a later even one
    x[4, :, 4 % V] = x[4, :, 1] = 50.0
    x[5, 5:9, :] = -100.0                         # -100 rows -> pad when the fix-up is on
    x[6, :, V - 1] = np.inf
    lengths = rng.integers(1, T + 1, size=B).astype(np.int32)
    lengths[:7] = T
    flat = torch.zeros(B * T * V + 8, dtype=torch.float32, device="cuda")
    view = flat[offset: offset + B * T * V].view(B, T, V)
    view.copy_(torch.from_numpy(x))
    d_len = torch.from_numpy(lengths).cuda()
    for fix in (0, 1):
        ids = torch.full((B, T), -1, dtype=torch.int32, device="cuda")
        tok = torch.zeros((B, T), dtype=torch.int32, device="cuda")
        ln = torch.zeros(B, dtype=torch.int32, device="cuda")
        _lib.check(_lib.load().coral_ctc_greedy(view.data_ptr(), d_len.data_ptr(), B, T, V, V - 1, fix, ids.data_ptr(),
                                                tok.data_ptr(), ln.data_ptr(), _lib.stream_ptr(view.device)))
        got = ids.cpu().numpy()
        ref = np.argmax(x, axis=-1)
        if fix:
            ref = ref.copy()
            ref[np.all(x == -100.0, axis=-1)] = V - 1
        for b in range(B):
            assert np.array_equal(got[b, : lengths[b]], ref[b, : lengths[b]]), (b, fix)


def test_compute_error_rate_metrics_matches_oracle(torch_cuda, rng):
    from types import SimpleNamespace

    from coral_b200.compute_metrics import compute_error_rate_metrics
    from oracle import greedy as og

    B, T, V = 16, 120, 46
    preds = (rng.standard_normal((B, T, V)) * 3).astype(np.float32)
    preds[:, 80:, :] = -100.0  # Trainer pads the time axis with -100 (HF:trainer.py:2678-2722)
    preds[5, 40:, :] = -100.0
    labels = rng.integers(0, 42, size=(B, 30)).astype(np.int64)
    labels[:, 20:] = -100
    labels[:, 5] = 36
    labels[:, 12] = 36

    class Tok:
        pad_token_id = 45
        word_delimiter_token = "|"
        unk_token = "<unk>"
        do_lower_case = False

        def get_vocab(self):
            return {t: i for i, t in enumerate(og.CORAL_VOCAB)}

    ref = og.compute_error_rate_metrics(preds, labels)
    lab2 = labels.copy()
    got = compute_error_rate_metrics(SimpleNamespace(predictions=preds.copy(), label_ids=lab2),
                                     SimpleNamespace(tokenizer=Tok()), log_examples=False)
    assert got == ref
    assert (lab2 != -100).all()  # same in-place fix-up as R:src/coral/compute_metrics.py:46-48


# -------------------------------------------------------------------------------- edit
def _random_pairs(rng, n, alphabet, max_len, p_ws=0.15):
    refs, hyps = [], []
    for _ in range(n):
        L = int(rng.integers(1, max_len))
        ref = "".join(rng.choice(list(alphabet), size=L))
        if not ref.strip():
            ref = "a" + ref
        h = list(ref)
        for _ in range(int(rng.integers(0, max(2, L // 4)))):
            k = int(rng.integers(0, 3))
            pos = int(rng.integers(0, len(h) + 1))
            if k == 0 and h:
                h[min(pos, len(h) - 1)] = str(rng.choice(list(alphabet)))
            elif k == 1 and h:
                del h[min(pos, len(h) - 1)]
            else:
                h.insert(pos, str(rng.choice(list(alphabet))))
        refs.append(ref)
        hyps.append("".join(h))
    return refs, hyps


@pytest.mark.parametrize("max_len", [40, 128, 300, 700])
def test_edit_counts_bit_exact(torch_cuda, rng, max_len):
    from coral_b200.metrics import edit_counts
    from oracle import edit as oe

    n = 400 if max_len <= 128 else 60
    refs, hyps = _random_pairs(rng, n, "abcdeæøå  \t", max_len)
    hyps[0] = ""                      # empty hypothesis: all deletions
    hyps[1] = refs[1]                 # identical
    refs[2], hyps[2] = "ab", "ba"     # SURVEY known answer (0, 1, 1)
    refs[3], hyps[3] = "hej med dig", "hej  med   dig"
    cc = edit_counts(hyps, refs, "chars")
    wc = edit_counts(hyps, refs, "words")
    for i in range(n):
        assert tuple(cc[i]) == oe.char_counts(refs[i], hyps[i]), (i, refs[i], hyps[i])
        assert tuple(wc[i]) == oe.word_counts(refs[i], hyps[i]), (i, refs[i], hyps[i])
    assert tuple(cc[2][:3]) == (0, 1, 1)


def test_cer_wer_and_errors(torch_cuda, rng):
    from coral_b200 import metrics
    from oracle import edit as oe

    refs, hyps = _random_pairs(rng, 300, "abcdefg hij", 60)
    for norm in (True, False):
        assert metrics.cer(hyps, refs, norm) == oe.cer(hyps, refs, norm)
        assert metrics.wer(hyps, refs, norm) == oe.wer(hyps, refs, norm)
    # an empty reference scores as all insertions (jiwer >= 3.1; the reference pins 4.0.0) ...
    assert metrics.edit_counts(["abc", "x y"], ["", "   "], "chars").tolist() == [[0, 0, 3, 0], [0, 0, 3, 0]]
    assert metrics.edit_counts(["abc", "x y"], ["", "   "], "words").tolist() == [[0, 0, 1, 0], [0, 0, 2, 0]]
    assert metrics.cer(["abc", "de"], ["", "dx"]) == oe.cer(["abc", "de"], ["", "dx"]) == 4 / 5
    with pytest.raises(ZeroDivisionError):   # ... and only an empty total divides by zero
        metrics.cer(["abc"], [""], normalise=False)
    with pytest.raises(ZeroDivisionError):
        metrics.cer([], [])
    # the jiwer 3.0.x behaviour is one switch away
    import importlib
    os.environ["CORAL_B200_EMPTY_REFERENCE"] = "raise"
    try:
        importlib.reload(metrics)
        with pytest.raises(ValueError):
            metrics.cer(["abc"], [""])
        with pytest.raises(ValueError):
            metrics.wer(["abc"], ["   "])
    finally:
        del os.environ["CORAL_B200_EMPTY_REFERENCE"]
        importlib.reload(metrics)


def test_edit_counts_capacity_is_reported_not_truncated(torch_cuda, rng):
    """A string longer than the declared max_len is an error status, never a silently truncated
    score; pairs beyond rapidfuzz's direct-alignment range are refused; launches on different
    streams use their own off-chip work areas."""
    from coral_b200 import _lib, metrics
    from coral_b200.textio import encode_utf32
    from oracle import edit as oe

    torch = torch_cuda
    refs = ["ab" * 150 + "b", "kort", "x" * 140]   # pair 0: nothing in common at either end, 301 symbols
    hyps = ["ba" * 145 + "c", "kart", "x" * 150]
    r_cps, r_off = encode_utf32(refs)
    h_cps, h_off = encode_utf32(hyps)
    d = lambda a: torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a).cuda()
    sdih, status = metrics.edit_counts_device(d(r_cps), d(r_off), d(h_cps), d(h_off), 3, 1, 160)  # under-declared
    assert status.cpu().tolist() == [2, 0, 0] and sdih.cpu().tolist()[0] == [0, 0, 0, 0]
    assert tuple(sdih.cpu().tolist()[1]) == oe.char_counts(refs[1], hyps[1])
    with pytest.raises(_lib.CoralError):
        metrics.edit_counts(["a" * 2100], ["b" * 2100], "chars")
    # explicit spans: one that ends before it begins is refused (status 2), its neighbours are scored
    beg = torch.tensor([0, 5, 9], dtype=torch.int64).cuda()
    end = torch.tensor([4, 3, 12], dtype=torch.int64).cuda()
    cps = d(encode_utf32(["kortkartabcx"])[0])
    for mode, max_len in ((1, 16), (1, 600), (0, 16)):
        sdih, status = metrics.edit_counts_spans_device(cps, beg, end, cps, beg, end, 3, mode, max_len)
        assert status.cpu().tolist() == [0, 2, 0] and sdih.cpu().tolist() == [[0, 0, 0, 4], [0, 0, 0, 0], [0, 0, 0, 3]]
    # two streams, long strings (off-chip work areas), interleaved launches
    n = 64
    big_r = ["".join(rng.choice(list("abcdef "), size=int(rng.integers(200, 900)))) for _ in range(n)]
    big_h = ["".join(rng.choice(list("abcdef "), size=int(rng.integers(200, 900)))) for _ in range(n)]
    want = [oe.char_counts(r, h) for r, h in zip(big_r, big_h)]
    bufs = [tuple(d(x) for x in (*encode_utf32(big_r), *encode_utf32(big_h))) for _ in range(2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    outs = []
    for rep in range(3):
        for k, st in enumerate(streams):
            with torch.cuda.stream(st):
                rc, ro, hc, ho = bufs[k]
                outs.append(metrics.edit_counts_device(rc, ro, hc, ho, n, 1, 900))
    torch.cuda.synchronize()
    for sdih, status in outs:
        assert [tuple(x) for x in sdih.cpu().tolist()] == want and not status.cpu().numpy().any()


def test_edit_counts_two_kernels_agree(torch_cuda, rng):
    """The bit-parallel kernel and the general anti-diagonal kernel give the same counts (and both the
    oracle's), pair by pair, in all three modes; the token 0xffffffff (the bit-parallel kernel's
    empty-slot marker) is handed to the general kernel, not mis-scored."""
    import subprocess, sys, json
    from coral_b200 import metrics
    from coral_b200.textio import encode_utf32
    from oracle import edit as oe

    torch = torch_cuda
    refs, hyps = _random_pairs(rng, 3000, "abcdeæøå  \t", 250)
    for k in range(0, 3000, 7):        # unrelated pairs: cores as long as the strings
        hyps[k] = "".join(rng.choice(list("abcde æ"), size=int(rng.integers(1, 250))))
    # every white-space code point of str.isspace() / regex \\s, and their neighbours that are not
    spaces = [9, 10, 11, 12, 13, 28, 29, 30, 31, 32, 133, 160, 5760, 8232, 8233, 8239, 8287, 12288, *range(8192, 8203)]
    near = [8, 14, 27, 33, 63, 64, 132, 134, 159, 161, 5759, 5761, 8191, 8203, 8231, 8234, 8238, 8240, 8286, 8288,
            12287, 12289]
    for k, cp in enumerate(spaces + near):
        ch = chr(cp)
        refs[10 + 5 * k] = f"{ch}ab{ch}cd{ch}{ch}ef g{ch} h{ch}"
        hyps[10 + 5 * k] = f"ab cd{ch}ef  g h{ch}{ch}i"
    got_c = metrics.edit_counts(hyps, refs, "chars")
    got_w = metrics.edit_counts(hyps, refs, "words")
    code = (
        "import os, sys, json; os.environ['CORAL_B200_EDIT_NO_BITPAR'] = '1'; sys.path.insert(0, %r)\n"
        "from coral_b200 import metrics\n"
        "refs, hyps = json.load(sys.stdin)\n"
        "print(json.dumps([metrics.edit_counts(hyps, refs, k).tolist() for k in ('chars', 'words')]))\n"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], input=json.dumps([refs, hyps]), capture_output=True,
                         text=True, check=True).stdout
    old_c, old_w = json.loads(out.strip().splitlines()[-1])
    assert got_c.tolist() == old_c and got_w.tolist() == old_w
    for i in range(0, 3000, 5):
        assert tuple(got_c[i]) == oe.char_counts(refs[i], hyps[i])
        assert tuple(got_w[i]) == oe.word_counts(refs[i], hyps[i])
    # tokens mode
    n = 500
    a = [rng.integers(0, 6, size=int(rng.integers(0, 200))).astype(np.uint32) for _ in range(n)]
    b = [x.copy() if rng.random() < 0.5 else rng.integers(0, 6, size=int(rng.integers(0, 200))).astype(np.uint32)
         for x in a]
    for k in range(0, n, 3):
        if len(a[k]):
            a[k][rng.integers(0, len(a[k]))] = 0xFFFFFFFF
        if len(b[k]) > 3:
            b[k][rng.integers(0, len(b[k]))] = 0xFFFFFFFF
            b[k] = np.delete(b[k], 1)
    def flat(xs):
        off = np.zeros(len(xs) + 1, np.int64)
        off[1:] = np.cumsum([len(x) for x in xs])
        cps = np.concatenate(xs + [np.zeros(1, np.uint32)])
        return torch.from_numpy(cps.view(np.int32)).cuda(), torch.from_numpy(off).cuda()
    (rc, ro), (hc, ho) = flat(a), flat(b)
    sdih, status = metrics.edit_counts_device(rc, ro, hc, ho, n, 0, 200)
    sdih, status = sdih.cpu().numpy(), status.cpu().numpy()
    for i in range(n):
        S, D, I = oe.editops_counts(a[i].tolist(), b[i].tolist())
        assert tuple(sdih[i]) == (S, D, I, len(a[i]) - S - D), i
        assert status[i] == (1 if len(a[i]) == 0 else 0)


def test_get_score_df_and_validation(torch_cuda, rng):
    import pandas as pd

    from coral_b200.evaluate import get_score_df
    from coral_b200.validation import validation_scores
    from oracle import edit as oe

    refs, hyps = _random_pairs(rng, 240, "abcdefg hij", 50)
    df = pd.DataFrame(dict(
        age_group=rng.choice(["0-25", "25-50", "50+"], size=240),
        gender=rng.choice(["female", "male"], size=240),
        dialect=rng.choice(["a", "b", "c", "d"], size=240),
        prediction=hyps, text=refs))
    got = get_score_df(df, ["age_group", "gender", "dialect"])
    ref = oe.get_score_records(df.to_dict("records"), ["age_group", "gender", "dialect"])
    # same records, same order; exact float equality (integer counts -> one division)
    pd.testing.assert_frame_equal(got, pd.DataFrame.from_records(ref), check_exact=True)
    vs = validation_scores(hyps, refs, max_cer=0.6)
    assert vs.cer == oe.cer(hyps, refs) and vs.wer == oe.wer(hyps, refs)
    assert np.array_equal(vs.asr_cer, np.array(oe.per_sample_rates(hyps, refs, True, "cer")))
    assert np.array_equal(vs.keep, vs.asr_cer < 0.6)


# ---------------------------------------------------------------------------------- LM
def test_device_lm_scores_bit_exact(gpu_decoder, oracle_decoder, small_lm, rng):
    import synth

    words, model, _ = small_lm
    flat, lens = model.sample(300, "lmtest")
    sents = [s.split(" ") for s in synth.sentences_to_text(flat, lens, words)]
    for s in sents[::3]:
        s[int(rng.integers(len(s)))] = "zzzoov"
    km = gpu_decoder._language_model.kenlm_model
    m = oracle_decoder._language_model._kenlm_model
    assert km.order == m.order and ("zzzoov" in km) is False and (words[0] in km) == (words[0] in m)
    for bos, eos in ((True, True), (False, False)):
        probs, oovs = km.score_sentences(sents, bos=bos, eos=eos)
        for s, p, o in zip(sents, probs, oovs):
            st = m.begin_sentence_state() if bos else m.null_context_state()
            ref = []
            for w in s:
                x, st = m.base_score(st, w)
                ref.append(x)
            if eos:
                ref.append(m.base_score(st, "</s>")[0])
            assert np.array_equal(np.array(ref, dtype=np.float32), p)
            assert [int(w not in m) for w in s] == o.tolist()


# -------------------------------------------------------------------------------- beam
def test_beam_search_peaky_all_beams(gpu_decoder, oracle_decoder, small_workload):
    w = small_workload
    lg = [w.logits[u, : w.lengths[u]] for u in range(len(w.lengths))]
    got = gpu_decoder.decode_beams_batch(None, lg)
    for u in range(len(lg)):
        ref = oracle_decoder.decode_beams(lg[u])
        beams_equal(ref, got[u])
    assert gpu_decoder.decode_batch(None, lg) == [oracle_decoder.decode(x) for x in lg]
    assert gpu_decoder.decode(lg[0]) == oracle_decoder.decode(lg[0])
    b5 = gpu_decoder.decode_beams(lg[1])
    assert len(b5[0]) == 5 and b5[0][0] == oracle_decoder.decode(lg[1])


@pytest.mark.parametrize("beam_width,T", [(16, 80), (100, 60), (200, 40), (512, 24)])
def test_beam_search_flat_logits_trim_path(gpu_decoder, oracle_decoder, rng, beam_width, T):
    import synth

    lg = [synth.flat_logits(T, rng), synth.flat_logits(max(1, T // 2), rng)]
    got = gpu_decoder.decode_beams_batch(None, lg, beam_width=beam_width)
    for x, g in zip(lg, got):
        beams_equal(oracle_decoder.decode_beams(x, beam_width=beam_width), g)


def test_beam_search_parameters(gpu_decoder, oracle_decoder, small_workload):
    w = small_workload
    lg = [w.logits[u, : w.lengths[u]] for u in range(4)]
    cases = [
        dict(beam_width=25, beam_prune_logp=-5.0, token_min_logp=-3.0),
        dict(beam_width=64, beam_prune_logp=-20.0, token_min_logp=-10.0),
        dict(beam_width=128, beam_prune_logp=-10.0, token_min_logp=-20.0),
    ]
    for kw in cases:
        got = gpu_decoder.decode_beams_batch(None, lg, **kw)
        for x, g in zip(lg, got):
            beams_equal(oracle_decoder.decode_beams(x, **kw), g)
    # probabilities as input in the narrow instantiations (their staging scratch is the tightest)
    import math
    z = lg[0].astype(np.float64)
    p = np.exp(z - z.max(axis=1, keepdims=True))
    p = (p / p.sum(axis=1, keepdims=True)).astype(np.float32)
    if math.isclose(float(p.sum(axis=1).mean()), 1):
        for bw in (16, 48):
            beams_equal(oracle_decoder.decode_beams(p, beam_width=bw), gpu_decoder.decode_beams_batch(None, [p], beam_width=bw)[0])
    try:
        for dec in (gpu_decoder, oracle_decoder):
            dec.reset_params(alpha=0.9, beta=0.3, unk_score_offset=-4.0, lm_score_boundary=False)
        got = gpu_decoder.decode_beams_batch(None, lg)
        for x, g in zip(lg, got):
            beams_equal(oracle_decoder.decode_beams(x), g)
    finally:
        for dec in (gpu_decoder, oracle_decoder):
            dec.reset_params(alpha=0.5, beta=1.5, unk_score_offset=-10.0, lm_score_boundary=True)


def test_beam_search_no_lm_edge_cases_and_errors(torch_cuda, small_workload, rng):
    import synth
    from coral_b200.decoder import build_ctcdecoder
    from oracle.beam import build_ctcdecoder as oracle_build

    g = build_ctcdecoder(synth.CORAL_LABELS)
    o = oracle_build(synth.CORAL_LABELS)
    w = small_workload
    lg = [w.logits[0, : w.lengths[0]], synth.flat_logits(30, rng), w.logits[1, :1],
          np.zeros((0, 46), np.float32)]
    got = g.decode_beams_batch(None, lg)
    for x, gb in zip(lg, got):
        beams_equal(o.decode_beams(x), gb)
    # probabilities instead of logits (pyctcdecode's auto-detection)
    z = w.logits[2, : w.lengths[2]].astype(np.float64)
    p = np.exp(z - z.max(axis=1, keepdims=True))
    p = (p / p.sum(axis=1, keepdims=True)).astype(np.float32)
    import math
    if math.isclose(float(p.sum(axis=1).mean()), 1):
        beams_equal(o.decode_beams(p), g.decode_beams_batch(None, [p])[0])
    # a multi-code-point label ("<s>", id 42) that wins: decode_batch must fall back from the
    # device-side text path and still spell it out
    sp = np.full((12, 46), -8.0, np.float32)
    sp[:, 45] = 0.0
    sp[2, 42] = 6.0
    sp[5, 17] = 6.0
    sp[8, 43] = 6.0
    assert g.decode_batch(None, [sp, lg[0]]) == [o.decode(sp), o.decode(lg[0])]
    assert "<s>" in o.decode(sp)
    with pytest.raises(ValueError):
        g.decode_beams(np.zeros((10, 40), np.float32))
    with pytest.raises(ValueError):
        g.decode_beams(np.zeros((2, 10, 46), np.float32))
    with pytest.raises(NotImplementedError):
        g.decode_beams(lg[0], hotwords=["hej"])


def test_prune_history(gpu_decoder, oracle_decoder, small_workload, rng):
    """pyctcdecode's prune_history=True on the device: every beam, text and word frames, == the oracle;
    the text-only kernel agrees; without an LM the history is one word."""
    import synth
    from coral_b200.decoder import build_ctcdecoder
    from oracle.beam import build_ctcdecoder as oracle_build

    w = small_workload
    lg = [w.logits[u, : w.lengths[u]] for u in range(8)] + [synth.flat_logits(40, rng)]
    for bw in (100, 32, 200):
        got = gpu_decoder.decode_beams_batch(None, lg, beam_width=bw, prune_history=True)
        for x, g in zip(lg, got):
            ref = oracle_decoder.decode_beams(x, beam_width=bw, prune_history=True)
            beams_equal(ref, g)
    one = gpu_decoder.decode_beams(lg[0], prune_history=True)
    beams_equal(oracle_decoder.decode_beams(lg[0], prune_history=True), one)
    pad = np.full((len(lg), max(x.shape[0] for x in lg), 46), -100.0, np.float32)
    for i, x in enumerate(lg):
        pad[i, : x.shape[0]] = x
    out = gpu_decoder.decode_padded(pad, np.array([x.shape[0] for x in lg], np.int32), prune_history=True, n_best=100)
    texts = gpu_decoder.tokens_to_text(out.tokens.reshape(len(lg) * 100, -1), out.lens.reshape(-1))
    for u, x in enumerate(lg):
        ref = oracle_decoder.decode_beams(x, prune_history=True)
        g = [(texts[u * 100 + k], float(out.logit_score[u, k]), float(out.lm_score[u, k])) for k in range(int(out.n_beams[u]))]
        beams_equal(ref, g)
    g0, o0 = build_ctcdecoder(synth.CORAL_LABELS), oracle_build(synth.CORAL_LABELS)
    for x in (lg[0], lg[-1]):
        beams_equal(o0.decode_beams(x, prune_history=True), g0.decode_beams(x, prune_history=True))


def test_hf_processor_with_lm_drop_in(gpu_decoder, oracle_decoder, small_workload, tmp_path):
    """The unmodified HF Wav2Vec2ProcessorWithLM drives the CUDA decoder through the shims."""
    import json

    import multiprocessing

    import coral_b200

    coral_b200.install_shims()
    # HF forks a Pool around decode_beams_batch when the start method is "fork"
    # (HF:...processing_wav2vec2_with_lm.py:374-389); with "spawn" it decodes in-process.
    multiprocessing.set_start_method("spawn", force=True)
    from transformers import Wav2Vec2CTCTokenizer, Wav2Vec2FeatureExtractor, Wav2Vec2ProcessorWithLM

    import synth

    vocab = {c: i for i, c in enumerate(synth.CORAL_LABELS[:42])}
    (tmp_path / "vocab.json").write_text(json.dumps(vocab))
    tok = Wav2Vec2CTCTokenizer(str(tmp_path / "vocab.json"), unk_token="<unk>", pad_token="<pad>",
                               bos_token="<s>", eos_token="</s>", word_delimiter_token="|")
    fe = Wav2Vec2FeatureExtractor(feature_size=1, sampling_rate=16000, padding_value=0.0, do_normalize=True)
    proc = Wav2Vec2ProcessorWithLM(feature_extractor=fe, tokenizer=tok, decoder=gpu_decoder)
    w = small_workload
    out = proc.batch_decode(w.logits[:6])  # padded with -100 rows; HF strips them (:371)
    ref = [oracle_decoder.decode_beams(w.logits[u, : w.lengths[u]])[0] for u in range(6)]
    assert out.text == [r[0] for r in ref]
    for a, r in zip(out.logit_score, ref):
        assert abs(a - r[3]) <= 1e-4 * max(1, abs(r[3]))
    one = proc.decode(w.logits[0, : w.lengths[0]])
    assert one.text == ref[0][0]
    # word offsets (HF:...processing_wav2vec2_with_lm.py:416-443) come from the decoder's text_frames
    wo = proc.batch_decode(w.logits[:6], output_word_offsets=True)
    for offs, r in zip(wo.word_offsets, ref):
        assert [(d["word"], (d["start_offset"], d["end_offset"])) for d in offs] == \
               [(wd, (int(a), int(b))) for wd, (a, b) in r[2]]
    # n_best > 1 and the LM weights HF forwards to reset_params (HF:...processing_wav2vec2_with_lm.py:365-367)
    nb = proc.batch_decode(w.logits[:4], n_best=3, alpha=0.8, beta=0.7, unk_score_offset=-6.0, lm_score_boundary=False)
    try:
        oracle_decoder.reset_params(alpha=0.8, beta=0.7, unk_score_offset=-6.0, lm_score_boundary=False)
        for u in range(4):
            rb = oracle_decoder.decode_beams(w.logits[u, : w.lengths[u]])[:3]
            assert list(nb.text[u]) == [b[0] for b in rb]
            for a, b in zip(nb.lm_score[u], rb):
                assert abs(a - b[4]) <= 1e-4 * max(1.0, abs(b[4]))
    finally:
        oracle_decoder.reset_params(alpha=0.5, beta=1.5, unk_score_offset=-10.0, lm_score_boundary=True)
        gpu_decoder.reset_params(alpha=0.5, beta=1.5, unk_score_offset=-10.0, lm_score_boundary=True)
    # round trip through pyctcdecode's directory layout
    proc.save_pretrained(str(tmp_path / "m"))
    assert (tmp_path / "m" / "alphabet.json").exists() and (tmp_path / "m" / "language_model" / "attrs.json").exists()
    proc2 = Wav2Vec2ProcessorWithLM.from_pretrained(str(tmp_path / "m"))
    assert proc2.batch_decode(w.logits[:3]).text == out.text[:3]


def test_full_size_properties(gpu_decoder, torch_cuda, cache_dir, rng):
    """Size-independent properties at a larger batch: determinism across launches and slot
    reuse, n_best=1 equals the head of the full beam list, identical inputs give identical
    outputs wherever they sit in the batch."""
    import synth

    w = synth.build_workload(cache_dir, 256, order=4, n_words=2000, n_sent=5000, name="big")
    a = gpu_decoder.decode_padded(w.logits, w.lengths, n_best=1, collect_stats=True)
    b = gpu_decoder.decode_padded(w.logits, w.lengths, n_best=1)
    assert np.array_equal(a.tokens, b.tokens) and np.array_equal(a.logit_score, b.logit_score)
    assert a.stats[3] == int(w.lengths.sum())  # every frame was processed exactly once
    perm = rng.permutation(256)
    c = gpu_decoder.decode_padded(w.logits[perm], w.lengths[perm], n_best=1)
    assert np.array_equal(c.lens[:, 0], a.lens[perm, 0]) and np.array_equal(c.lm_score, a.lm_score[perm])
    full = gpu_decoder.decode_padded(w.logits[:32], w.lengths[:32], n_best=100)
    assert np.array_equal(full.lm_score[:, 0], a.lm_score[:32, 0])
    assert (np.diff(full.lm_score, axis=1)[np.arange(100)[None, 1:] < full.n_beams[:, None]] <= 0).all()


def test_streamed_host_input_equals_resident_input(gpu_decoder, oracle_decoder, torch_cuda, cache_dir, rng):
    """Pinned host logits take the single-launch streamed path (chunks + ready counter): same
    beams as the device-resident launch, ragged last chunk included; and the probabilities-
    vs-logits detection is per utterance inside the kernel (a mixed batch)."""
    import math

    import synth
    from coral_b200 import decoder as dmod

    torch = torch_cuda
    w = synth.build_workload(cache_dir, 256, order=4, n_words=2000, n_sent=5000, name="big")
    idx = np.concatenate([rng.permutation(256) for _ in range(5)])[: 2 * dmod.H2D_CHUNK + 173]
    lg = torch.from_numpy(w.logits[idx]).pin_memory()
    ln = w.lengths[idx]
    resident = gpu_decoder.decode_padded(lg.cuda(), ln, n_best=3)
    for _ in range(2):
        streamed = gpu_decoder.decode_padded(lg, ln, n_best=3)
        for k in ("n_beams", "tokens", "lens"):
            assert np.array_equal(getattr(streamed, k), getattr(resident, k)), k
        valid = np.arange(3)[None, :] < resident.n_beams[:, None]  # score slots past n_beams are not written
        for k in ("logit_score", "lm_score"):
            assert np.array_equal(getattr(streamed, k)[valid], getattr(resident, k)[valid]), k
    # mixed batch: every third utterance as probabilities
    mixed = [w.logits[u, : w.lengths[u]] for u in range(12)]
    for u in range(0, 12, 3):
        z = mixed[u].astype(np.float64)
        p = np.exp(z - z.max(axis=1, keepdims=True))
        p = (p / p.sum(axis=1, keepdims=True)).astype(np.float32)
        if math.isclose(float(p.sum(axis=1).mean()), 1):
            mixed[u] = p
    got = gpu_decoder.decode_beams_batch(None, mixed)
    for x, g in zip(mixed, got):
        beams_equal(oracle_decoder.decode_beams(x), g)


def test_decode_batches_one_ahead_equals_batch_by_batch(gpu_decoder, torch_cuda, cache_dir, rng):
    """decode_batches (the next batch decoding while the caller scores this one) yields what
    decode_batch returns batch by batch, in order, for every input form; the yielded lists carry
    their device text, and cer()/wer() queued by the caller between batches are right."""
    import synth
    from coral_b200 import metrics
    from oracle import edit as oe

    torch = torch_cuda
    w = synth.build_workload(cache_dir, 256, order=4, n_words=2000, n_sent=5000, name="big")
    idxs = [rng.permutation(256)[: int(n)] for n in (200, 1, 64, 256, 17)]
    want = [list(gpu_decoder.decode_batch(None, torch.from_numpy(w.logits[i]).cuda(), lengths=w.lengths[i]))
            for i in idxs]

    def forms():
        for k, i in enumerate(idxs):
            if k % 3 == 0:
                yield (torch.from_numpy(w.logits[i]).pin_memory(), w.lengths[i])        # pinned: read in place
            elif k % 3 == 1:
                yield [w.logits[u, : w.lengths[u]] for u in i]                           # list of [T_i, V] arrays
            else:
                yield (torch.from_numpy(w.logits[i]).cuda(), torch.from_numpy(w.lengths[i]))  # device resident

    for prefetch in (1, 0, 3):
        seen = 0
        for k, got in enumerate(gpu_decoder.decode_batches(forms(), prefetch=prefetch)):
            assert list(got) == want[k], (prefetch, k)
            refs = [w.references[u] for u in idxs[k]]
            assert getattr(got, "_coral_dev", None) is not None
            assert metrics.cer(got, refs) == oe.cer(want[k], refs) and metrics.wer(got, refs) == oe.wer(want[k], refs)
            seen += 1
        assert seen == len(idxs)
    # a caller that stops early: the decodes still in flight are waited for, later calls are unaffected
    it = gpu_decoder.decode_batches(forms())
    assert list(next(it)) == want[0]
    it.close()
    assert list(gpu_decoder.decode_batch(None, [w.logits[u, : w.lengths[u]] for u in idxs[2]])) == want[2]
    assert list(gpu_decoder.decode_batches([])) == []
    assert [list(x) for x in gpu_decoder.decode_batches([[], [w.logits[3, : w.lengths[3]]]])] == \
        [[], list(gpu_decoder.decode_batch(None, [w.logits[3, : w.lengths[3]]]))]


def test_host_inputs_are_read_in_place_and_packed_ragged(gpu_decoder, torch_cuda, cache_dir, rng):
    """Host logits never get a staging copy on the device: a pinned padded tensor is read in place
    by the kernel, pageable arrays (a padded array, or the list of [T_i, V] arrays HF hands over)
    are packed into ONE pinned ragged buffer chunk by chunk while the kernel runs. Different
    batches alternate through the same staging buffer (a stale cached line of the previous batch
    would show), the transcripts carry their device-resident text into cer()/wer(), and all of it
    equals the device-resident launch."""
    import synth
    from coral_b200 import decoder as dmod, metrics
    from oracle import edit as oe

    torch = torch_cuda
    w = synth.build_workload(cache_dir, 256, order=4, n_words=2000, n_sent=5000, name="big")
    n = 2 * dmod.H2D_CHUNK + 91
    batches = []
    for _ in range(2):
        idx = np.concatenate([rng.permutation(256) for _ in range(6)])[:n]
        want = gpu_decoder.decode_padded(torch.from_numpy(w.logits[idx]).cuda(), w.lengths[idx], n_best=1)
        texts = gpu_decoder.tokens_to_text(want.tokens[:, 0, :], want.lens[:, 0])
        batches.append((idx, want, texts))
    for rep in range(3):
        for idx, want, texts in batches:
            lst = [w.logits[u, : w.lengths[u]] for u in idx]           # pageable, ragged
            got = gpu_decoder.decode_batch(None, lst)
            assert list(got) == texts
            refs = [w.references[u] for u in idx]
            assert getattr(got, "_coral_dev", None) is not None       # text stayed on the device
            assert metrics.cer(got, refs) == oe.cer(texts, refs) and metrics.wer(got, refs) == oe.wer(texts, refs)
            if rep == 0:
                pinned = torch.from_numpy(w.logits[idx]).pin_memory()  # pinned, padded: zero-copy
                a = gpu_decoder.decode_padded(pinned, w.lengths[idx], n_best=1)
                b = gpu_decoder.decode_padded(w.logits[idx], w.lengths[idx], n_best=1)  # pageable, padded
                for k in ("n_beams", "tokens", "lens", "logit_score", "lm_score"):
                    assert np.array_equal(getattr(a, k), getattr(want, k)), k
                    assert np.array_equal(getattr(b, k), getattr(want, k)), k
    # a mutated transcript list must not be scored from the stale device copy
    got = gpu_decoder.decode_batch(None, [w.logits[u, : w.lengths[u]] for u in range(8)])
    refs = [w.references[u] for u in range(8)]
    got[0] = got[0] + " x"
    assert metrics.cer(got, refs) == oe.cer(list(got), refs)
    # small batches (no chunking) and a mixed-dtype / non-contiguous list
    odd = [np.asfortranarray(w.logits[0, : w.lengths[0]]), w.logits[1, : w.lengths[1]].astype(np.float64)]
    assert list(gpu_decoder.decode_batch(None, odd)) == list(gpu_decoder.decode_batch(None, [w.logits[0, : w.lengths[0]], w.logits[1, : w.lengths[1]]]))


def test_stalled_input_stream_fails_the_launch_instead_of_hanging(gpu_decoder, torch_cuda, small_workload, monkeypatch):
    """A ready counter that never moves (a copier that died) must end the launch with an error
    status for every utterance after ONE timeout, not hang the device."""
    import time

    torch = torch_cuda
    monkeypatch.setenv("CORAL_READY_TIMEOUT_CYCLES", str(400_000_000))  # ~0.2 s
    w = small_workload
    d_logits = torch.from_numpy(w.logits).cuda()
    d_len = torch.from_numpy(w.lengths).cuda()
    ready = torch.zeros(1, dtype=torch.int32, device="cuda")
    t0 = time.perf_counter()
    outs = gpu_decoder.decode_launch(d_logits, d_len, None, n_best=1, ready=(ready, 4))
    torch.cuda.synchronize()
    assert time.perf_counter() - t0 < 10.0
    status = outs[5].cpu().numpy()
    assert (status == -3).all() and (outs[0].cpu().numpy() == 0).all()
    monkeypatch.delenv("CORAL_READY_TIMEOUT_CYCLES")
    ok = gpu_decoder.decode_launch(d_logits, d_len, None, n_best=1)  # the decoder is fine afterwards
    torch.cuda.synchronize()
    assert int(ok[5].abs().sum().item()) == 0


def test_concurrent_streams_share_one_decoder(gpu_decoder, torch_cuda, cache_dir):
    """Launches of the same decoder handle on different CUDA streams may overlap: each stream
    has its own scratch arenas and work counter (SURVEY 8b threading contract)."""
    import synth

    torch = torch_cuda
    w = synth.build_workload(cache_dir, 256, order=4, n_words=2000, n_sent=5000, name="big")
    d_logits = torch.from_numpy(w.logits).cuda()
    d_len = torch.from_numpy(w.lengths).cuda()
    want = gpu_decoder.decode_padded(d_logits, d_len, n_best=2)
    streams = [torch.cuda.Stream() for _ in range(4)]
    torch.cuda.synchronize()
    outs = []
    for rep in range(3):
        for k, st in enumerate(streams):
            with torch.cuda.stream(st):
                a0, b0 = 64 * k, 64 * (k + 1)
                outs.append((a0, b0, gpu_decoder.decode_launch(d_logits[a0:b0], d_len[a0:b0], None, n_best=2)))
    torch.cuda.synchronize()
    for a0, b0, (d_n, d_logit, d_comb, d_tok, d_lens, d_status) in outs:
        assert int(d_status.sum().item()) == 0
        assert np.array_equal(d_n.cpu().numpy(), want.n_beams[a0:b0])
        assert np.array_equal(d_tok.cpu().numpy(), want.tokens[a0:b0])
        assert np.array_equal(d_comb.cpu().numpy()[:, 0], want.lm_score[a0:b0, 0])


def test_pipeline_on_device_handoff(gpu_decoder, oracle_decoder, torch_cuda, tmp_path, rng):
    """the ASR pipeline with ``pipeline_class=BatchedCTCWithLMPipeline`` decodes each
    model batch on the GPU where the logits are produced. Same texts and word timestamps as the
    stock pipeline driving the oracle decoder one utterance at a time (R:src/coral/evaluate.py:56-60)."""
    import json

    import coral_b200

    coral_b200.install_shims()
    from transformers import (AutomaticSpeechRecognitionPipeline, Wav2Vec2Config, Wav2Vec2CTCTokenizer,
                              Wav2Vec2FeatureExtractor, Wav2Vec2ForCTC, pipeline)

    import synth
    from coral_b200.pipeline import BatchedCTCWithLMPipeline

    torch = torch_cuda
    vocab = {c: i for i, c in enumerate(synth.CORAL_LABELS[:42])}
    (tmp_path / "vocab.json").write_text(json.dumps(vocab))
    tok = Wav2Vec2CTCTokenizer(str(tmp_path / "vocab.json"), unk_token="<unk>", pad_token="<pad>",
                               bos_token="<s>", eos_token="</s>", word_delimiter_token="|")
    fe = Wav2Vec2FeatureExtractor(feature_size=1, sampling_rate=16000, padding_value=0.0, do_normalize=True)
    torch.manual_seed(7)
    cfg = Wav2Vec2Config(vocab_size=46, hidden_size=32, num_hidden_layers=1, num_attention_heads=2,
                         intermediate_size=64, conv_dim=(16,) * 7, pad_token_id=45, num_conv_pos_embeddings=16,
                         num_conv_pos_embedding_groups=4)
    model = Wav2Vec2ForCTC(cfg).eval()
    with torch.no_grad():
        model.lm_head.weight.mul_(40.0)  # random-init heads are flat; make a few tokens stand out
    audios = [rng.standard_normal(int(16000 * d)).astype(np.float32) for d in (0.9, 1.4, 0.6, 1.1, 0.8)]
    ours = pipeline("automatic-speech-recognition", model=model, tokenizer=tok, feature_extractor=fe,
                    decoder=gpu_decoder, device=0, pipeline_class=BatchedCTCWithLMPipeline)
    stock = pipeline("automatic-speech-recognition", model=model, tokenizer=tok, feature_extractor=fe,
                     decoder=oracle_decoder, device=0)
    assert type(stock) is AutomaticSpeechRecognitionPipeline
    got = ours([a.copy() for a in audios], batch_size=4, return_timestamps="word")
    ref = stock([a.copy() for a in audios], batch_size=4, return_timestamps="word")
    assert [g["text"] for g in got] == [r["text"] for r in ref]
    assert [g["chunks"] for g in got] == [r["chunks"] for r in ref]
    one = ours(audios[1].copy(), decoder_kwargs={"beam_width": 20})
    assert one["text"] == stock(audios[1].copy(), decoder_kwargs={"beam_width": 20})["text"]


def test_kenlm_binary_model_directory(gpu_decoder, oracle_decoder, small_lm, small_workload, tmp_path):
    """a decoder directory in CoRal's shipped layout -- language_model/{N}gram.bin
    (KenLM probing binary) + unigrams.txt + attrs.json (R:src/coral/ngram.py:361-387) -- loads through
    load_from_dir and decodes exactly like the ARPA-built decoder."""
    import shutil
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from kenlm_binary_writer import write_probing_binary

    from coral_b200.decoder import BeamSearchDecoderCTC
    from oracle.arpa import ArpaModel

    d = tmp_path / "model"
    gpu_decoder.save_to_dir(str(d))
    lm_dir = d / "language_model"
    arpa = [f for f in os.listdir(lm_dir) if f.endswith(".arpa")]
    assert len(arpa) == 1
    write_probing_binary(ArpaModel.load(str(lm_dir / arpa[0])), str(lm_dir / "5gram.bin"))
    # a binary with its ARPA next to it (same stem) is cross-checked at load time -- and refused when
    # the two disagree (the reader is not pinned against a real build_binary file)
    from coral_b200 import _lib
    from coral_b200.language_model import KenlmModel

    side = tmp_path / "side"
    side.mkdir()
    shutil.copy(lm_dir / "5gram.bin", side / "m.bin")
    shutil.copy(lm_dir / arpa[0], side / "m.arpa")
    assert KenlmModel(str(side / "m.bin")).is_binary
    text = (side / "m.arpa").read_text().splitlines()
    k0 = next(i for i, ln in enumerate(text) if ln.startswith("\\1-grams:")) + 1
    k1 = next(i for i, ln in enumerate(text) if ln.startswith("\\2-grams:"))
    for k in range(k0, k1):  # every unigram probability moves: any sentence now scores differently
        f = text[k].split("\t")
        if len(f) >= 2:
            f[0] = "%.4f" % (float(f[0]) - 0.5)
            text[k] = "\t".join(f)
    (side / "m.arpa").write_text("\n".join(text) + "\n")
    with pytest.raises(_lib.CoralError):
        KenlmModel(str(side / "m.bin"))
    os.remove(lm_dir / arpa[0])
    dec = BeamSearchDecoderCTC.load_from_dir(str(d))
    assert dec._language_model._kenlm_model.order == gpu_decoder._language_model._kenlm_model.order
    w = small_workload
    lg = [w.logits[u, : w.lengths[u]] for u in range(6)]
    got = dec.decode_beams_batch(None, lg)
    for x, g in zip(lg, got):
        beams_equal(oracle_decoder.decode_beams(x), g)
    shutil.rmtree(d)


# ------------------------------------------------------------------- golden fixtures on GPU
def _golden(name):
    import json
    import os

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name), encoding="utf-8") as f:
        return json.load(f)


def test_golden_greedy_vectors_from_real_hf_tokenizer(torch_cuda, coral_vocab):
    """ids -> strings through the CUDA collapse kernel == the real Wav2Vec2CTCTokenizer."""
    from coral_b200.greedy import decode_ids

    cases = [c for c in _golden("greedy_hf.json") if "ids" in c and len(c["ids"]) > 0]
    T = max(len(c["ids"]) for c in cases)
    ids = np.full((len(cases), T), 45, dtype=np.int32)
    lens = np.array([len(c["ids"]) for c in cases], dtype=np.int32)
    for i, c in enumerate(cases):
        ids[i, : lens[i]] = c["ids"]
    assert decode_ids(ids, coral_vocab, lengths=lens, group_tokens=True) == [c["grouped"] for c in cases]
    assert decode_ids(ids, coral_vocab, lengths=lens, group_tokens=False) == [c["ungrouped"] for c in cases]


def test_golden_edit_lm_and_beam_anchors(torch_cuda):
    import os

    from coral_b200.decoder import build_ctcdecoder
    from coral_b200.metrics import edit_counts

    rows = _golden("edit_known.json")
    cc = edit_counts([r["hyp"] for r in rows], [r["ref"] for r in rows], "chars")
    wc = edit_counts([r["hyp"] for r in rows], [r["ref"] for r in rows], "words")
    for r, c, w in zip(rows, cc, wc):
        assert c.tolist() == r["chars"] and w.tolist() == r["words"], r

    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    data = _golden("beam_toy.json")
    logits = np.load(os.path.join(g, "beam_toy_logits.npz"))
    decs = {"lm": build_ctcdecoder(data["labels"], os.path.join(g, "toy.arpa")), "nolm": build_ctcdecoder(data["labels"])}
    km = decs["lm"]._language_model.kenlm_model
    for case in _golden("lm_toy.json"):
        probs, _ = km.score_sentences([case["sentence"]], bos=case["bos"], eos=True)
        assert probs[0].tolist() == [float(np.float32(x)) for x in case["log10"]]
    for case in data["cases"]:
        got = decs[case["decoder"]].decode_beams(logits[case["logits"]], **case["kwargs"])
        assert [b[0] for b in got] == [b[0] for b in case["beams"]]
        for b, ref in zip(got, case["beams"]):
            assert abs(b[3] - ref[1]) <= 1e-4 * max(1, abs(ref[1])) and abs(b[4] - ref[2]) <= 1e-4 * max(1, abs(ref[2]))


def test_config4_validation_properties_at_scale(torch_cuda, rng):
    """Config 4 shape (per-sample CER filter) at 50k pairs through size-independent properties:
    S + D + H == len(ref), I - D == len(hyp) - len(ref), identical pairs score zero, the
    aggregate equals the sum of the per-sample counts, and a 2k subsample is bit-exact."""
    import synth
    from coral_b200.metrics import edit_counts
    from coral_b200.validation import validation_scores
    from oracle import edit as oe

    words = synth.make_word_list(3000)
    n = 50_000
    idx = rng.integers(0, len(words), size=(n, 12))
    lens = rng.integers(3, 13, size=n)
    refs = [" ".join(words[j] for j in idx[i, : lens[i]]) for i in range(n)]
    hyps = [synth.corrupt_text(r, rng, 0.07) for r in refs]
    hyps[::1000] = refs[::1000]
    vs = validation_scores(hyps, refs, max_cer=0.6)
    cc, wc = vs.char_counts, vs.word_counts
    ref_len = np.array([len(r.strip()) for r in refs])
    hyp_len = np.array([len(h.strip()) for h in hyps])
    assert np.array_equal(cc[:, 0] + cc[:, 1] + cc[:, 3], ref_len)
    assert np.array_equal(cc[:, 2] - cc[:, 1], hyp_len - ref_len)
    assert (cc[::1000, :3] == 0).all() and (wc[::1000, :3] == 0).all()
    tot = cc.sum(axis=0)
    assert vs.cer == int(tot[0] + tot[1] + tot[2]) / int(tot[0] + tot[1] + tot[3] + tot[2])
    assert 0.03 < vs.cer < 0.12 and vs.keep.mean() > 0.99
    sub = rng.choice(n, size=2000, replace=False)
    for i in sub:
        assert tuple(cc[i]) == oe.char_counts(refs[i], hyps[i]) and tuple(wc[i]) == oe.word_counts(refs[i], hyps[i])
    assert np.array_equal(edit_counts(hyps[:64], refs[:64], "chars"), cc[:64])
