"""The oracle against the golden vectors in tests/golden/ (CPU).

``greedy_hf.json`` was produced by the REAL ``Wav2Vec2CTCTokenizer`` of the transformers
version the reference pins (tests/golden/make_golden.py), so it pins the greedy path to
the reference. The other files freeze the oracle's restatement of the un-vendored
packages (regression anchors + the hand-computed known answers of SURVEY.md section 8c)."""

from __future__ import annotations

import json
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    with open(os.path.join(G, name), encoding="utf-8") as f:
        return json.load(f)


def test_greedy_restatement_matches_real_hf_tokenizer_golden():
    from oracle.greedy import CORAL_PAD_ID, CORAL_VOCAB, ids_to_string

    data = _load("greedy_hf.json")
    meta = data[-1]
    assert meta["vocab"] == CORAL_VOCAB and meta["pad_id"] == CORAL_PAD_ID
    assert data[-2]["input_ids"] == [17, 14, 19, 36, 22, 14, 13, 36, 13, 18, 16]
    n = 0
    for case in data[:-2]:
        assert ids_to_string(case["ids"]) == case["grouped"]
        assert ids_to_string(case["ids"], group_tokens=False) == case["ungrouped"]
        n += 1
    assert n >= 100
    assert ids_to_string([45, 17, 17, 45, 17, 36, 36, 45, 36, 10, 44, 42, 37]) == "hh  a<unk><s>å"
    assert ids_to_string([45, 17, 17, 45, 17, 36, 36, 45, 36, 10, 44, 42, 37], group_tokens=False) == "hhh   a<unk><s>å"


def test_greedy_restatement_matches_installed_tokenizer_live(tmp_path):
    """Same check against the tokenizer installed here, when transformers is importable."""
    transformers = pytest.importorskip("transformers")
    from oracle.greedy import CORAL_CHARACTERS, ids_to_string

    vocab = {ch: i for i, ch in enumerate(sorted(set(CORAL_CHARACTERS + "|")))}
    (tmp_path / "vocab.json").write_text(json.dumps(vocab))
    tok = transformers.Wav2Vec2CTCTokenizer(str(tmp_path / "vocab.json"), unk_token="<unk>", pad_token="<pad>",
                                            bos_token="<s>", eos_token="</s>", word_delimiter_token="|")
    rng = np.random.default_rng(1)
    for _ in range(200):
        ids = np.repeat(rng.integers(0, 46, size=int(rng.integers(1, 40))), rng.integers(1, 3)).tolist()
        assert ids_to_string(ids) == tok.decode(ids)
        assert ids_to_string(ids, group_tokens=False) == tok.decode(ids, group_tokens=False)


def test_edit_known_answers():
    from oracle import edit

    rows = _load("edit_known.json")
    for r in rows:
        assert list(edit.char_counts(r["ref"], r["hyp"])) == r["chars"]
        assert list(edit.word_counts(r["ref"], r["hyp"])) == r["words"]
    by = {(r["ref"], r["hyp"]): r for r in rows}
    # hand-checked
    assert by[("ab", "ba")]["chars"][:3] == [0, 1, 1]
    assert by[("abc", "bcd")]["chars"][:3] == [0, 1, 1]
    assert by[("hej med dig", "")]["chars"] == [0, 11, 0, 0] and by[("hej med dig", "")]["words"] == [0, 3, 0, 0]
    assert by[("hej med dig", "hej  med   dig")]["words"] == [0, 0, 0, 3]   # WER collapses spaces
    assert by[("hej med dig", "hej  med   dig")]["chars"][2] == 3            # CER counts each one
    assert edit.cer(["ba"], ["ab"]) == 2 / 3 and edit.cer(["ba"], ["ab"], normalise=False) == 1.0
    # an empty reference: all insertions under jiwer >= 3.1 (the reference pins 4.0.0), ValueError under 3.0.x
    assert edit.char_counts("", "x y") == (0, 0, 3, 0) and edit.word_counts("  ", "x y") == (0, 0, 2, 0)
    assert edit.cer(["x"], [""]) == 1.0
    with pytest.raises(ZeroDivisionError):
        edit.cer(["x"], [""], normalise=False)
    edit.EMPTY_REFERENCE = "raise"
    try:
        with pytest.raises(ValueError):
            edit.cer(["x"], [""])
        with pytest.raises(ValueError):
            edit.wer(["x"], ["  "])
    finally:
        edit.EMPTY_REFERENCE = "allow"
    with pytest.raises(ZeroDivisionError):
        edit.cer([], [])


def test_lm_toy_scores_and_hand_computed_backoff():
    from oracle.arpa import ArpaModel, load_unigram_set_from_arpa

    m = ArpaModel.load(os.path.join(G, "toy.arpa"))
    assert m.order == 3 and "hej" in m and "ukendt" not in m and "<unk>" not in m
    for case in _load("lm_toy.json"):
        st = m.begin_sentence_state() if case["bos"] else m.null_context_state()
        got = []
        for w in case["sentence"].split():
            p, st = m.base_score(st, w)
            got.append(p)
        got.append(m.base_score(st, "</s>")[0])
        assert got == case["log10"]
    f32 = np.float32
    st = m.begin_sentence_state()
    p, st = m.base_score(st, "hej")          # bigram "<s> hej"
    assert p == float(f32(-0.4))
    p, st = m.base_score(st, "med")          # trigram "<s> hej med"
    assert p == float(f32(-0.2))
    p, st2 = m.base_score(st, "der")         # "hej med der" absent -> bigram "med der" + backoff("hej med")
    assert p == float(f32(f32(-1.1) + f32(-0.2)))
    p, _ = m.base_score(st, "ukendt")        # OOV -> <unk> unigram + backoff(med) + backoff(hej med)
    assert p == float(f32(f32(f32(-2.5) + f32(-0.3)) + f32(-0.2)))
    p, _ = m.base_score(m.null_context_state(), "hej")
    assert p == float(f32(-0.9))
    assert load_unigram_set_from_arpa(os.path.join(G, "toy.arpa")) == {
        "<unk>", "<s>", "</s>", "hej", "med", "dig", "der", "meget"}


def test_beam_toy_anchors():
    from oracle.beam import build_ctcdecoder

    data = _load("beam_toy.json")
    logits = np.load(os.path.join(G, "beam_toy_logits.npz"))
    decs = {"lm": build_ctcdecoder(data["labels"], os.path.join(G, "toy.arpa")), "nolm": build_ctcdecoder(data["labels"])}
    for case in data["cases"]:
        beams = decs[case["decoder"]].decode_beams(logits[case["logits"]], **case["kwargs"])
        assert [b[0] for b in beams] == [b[0] for b in case["beams"]]
        for b, ref in zip(beams, case["beams"]):
            assert abs(b[3] - ref[1]) < 1e-9 and abs(b[4] - ref[2]) < 1e-9
