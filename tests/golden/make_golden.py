"""Generates the golden fixtures under tests/golden/. Run from the repo root:

    python tests/golden/make_golden.py

* ``greedy_hf.json`` -- PINNED against real code: id sequences decoded by the real
  ``Wav2Vec2CTCTokenizer`` of transformers 5.5.0 (the version the reference pins,
  R:uv.lock:3290-3291), built exactly as R:src/coral/wav2vec2.py:64-72 builds it,
  with ``group_tokens`` on and off. The tokenizer is the reference's own greedy path
  (R:src/coral/compute_metrics.py:69-70), so these vectors pin oracle.greedy and the
  CUDA collapse kernel to the reference.
* ``toy.arpa`` + ``beam_toy.json`` / ``lm_toy.json`` / ``edit_known.json`` -- frozen outputs
  of the ORACLE (a restatement; pyctcdecode/kenlm/jiwer are not installable here, so these
  are regression anchors, not reference outputs: "parity unpinned", see oracle/__init__.py),
  plus the hand-computed known answers of SURVEY.md section 8c.
"""

from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CHARS = "abcdefghijklmnopqrstuvwxyzæøå0123456789éü"

TOY_ARPA = """\\data\\
ngram 1=8
ngram 2=7
ngram 3=3

\\1-grams:
-2.5\t<unk>\t0
-99\t<s>\t-0.5
-1.2\t</s>\t0
-0.9\thej\t-0.4
-1.0\tmed\t-0.3
-1.1\tdig\t-0.2
-1.6\tder\t-0.1
-1.9\tmeget\t0

\\2-grams:
-0.4\t<s> hej\t-0.25
-0.5\thej med\t-0.2
-0.3\tmed dig\t-0.15
-0.6\tdig </s>\t0
-0.8\thej der\t-0.1
-0.9\tder </s>\t0
-1.1\tmed der\t0

\\3-grams:
-0.2\t<s> hej med
-0.1\thej med dig
-0.15\tmed dig </s>

\\end\\
"""


def make_greedy():
    from transformers import Wav2Vec2CTCTokenizer

    vocab = {ch: i for i, ch in enumerate(sorted(set(CHARS + "|")))}
    d = tempfile.mkdtemp()
    json.dump(vocab, open(os.path.join(d, "vocab.json"), "w"))
    tok = Wav2Vec2CTCTokenizer(os.path.join(d, "vocab.json"), unk_token="<unk>", pad_token="<pad>",
                               bos_token="<s>", eos_token="</s>", word_delimiter_token="|",
                               replace_word_delimiter_char=" ")
    rng = np.random.default_rng(7)
    cases = [[45, 17, 17, 45, 17, 36, 36, 45, 36, 10, 44, 42, 37], [], [45, 45, 45], [36, 36, 10, 36, 36],
             [10, 10, 45, 10, 36, 45, 36, 11, 43, 43, 44]]
    for _ in range(120):
        L = int(rng.integers(1, 80))
        p = np.full(46, 0.3 / 45)
        p[45] = 0.45
        p[36] = 0.1
        p /= p.sum()
        ids = rng.choice(46, size=L, p=p)
        rep = rng.integers(1, 4, size=L)
        cases.append(np.repeat(ids, rep).tolist())
    out = []
    for ids in cases:
        out.append({
            "ids": ids,
            "grouped": tok.decode(ids),
            "ungrouped": tok.decode(ids, group_tokens=False),
        })
    out.append({"tokenize": "hej med dig", "input_ids": tok("hej med dig").input_ids})
    out.append({"vocab": [t for t, _ in sorted(tok.get_vocab().items(), key=lambda kv: kv[1])],
                "pad_id": tok.pad_token_id})
    json.dump(out, open(os.path.join(HERE, "greedy_hf.json"), "w"), ensure_ascii=False, indent=0)


def make_oracle_anchors():
    from oracle import edit
    from oracle.arpa import ArpaModel
    from oracle.beam import build_ctcdecoder

    arpa = os.path.join(HERE, "toy.arpa")
    open(arpa, "w", encoding="utf-8").write(TOY_ARPA)
    m = ArpaModel.load(arpa)
    lm_cases = []
    for sent in ["hej med dig", "hej der", "dig med hej", "hej ukendt dig", "meget meget", ""]:
        for bos in (True, False):
            st = m.begin_sentence_state() if bos else m.null_context_state()
            ps = []
            for w in sent.split():
                p, st = m.base_score(st, w)
                ps.append(p)
            ps.append(m.base_score(st, "</s>")[0])
            lm_cases.append({"sentence": sent, "bos": bos, "log10": ps})
    json.dump(lm_cases, open(os.path.join(HERE, "lm_toy.json"), "w"), indent=0)

    labels = ["a", "b", "d", "e", "g", "h", "i", "j", "m", "r", "t", "|", "<unk>", "<pad>"]
    rng = np.random.default_rng(11)
    dec = build_ctcdecoder(labels, arpa)
    dec0 = build_ctcdecoder(labels)
    L = dec._alphabet.labels

    def align(text, T):
        ids = [L.index(c) for c in text]
        lg = rng.standard_normal((T, len(L))).astype(np.float32)
        pos = np.sort(rng.choice(T, size=len(ids), replace=False))
        tgt = np.full(T, L.index(""))
        tgt[pos] = ids
        lg[np.arange(T), tgt] += 6.0
        alt = rng.integers(0, len(L), size=T)
        amb = rng.random(T) < 0.4
        lg[amb, alt[amb]] += 5.0
        return lg

    beam_cases = []
    logits_store = {}
    for k, (text, T) in enumerate([("hej med dig", 30), ("hej der", 20), ("meget med", 28), ("hjem", 9), ("", 5)]):
        lg = align(text, T)
        logits_store[f"logits_{k}"] = lg
        for name, d in (("lm", dec), ("nolm", dec0)):
            for kw in ({}, {"beam_width": 8, "beam_prune_logp": -6.0, "token_min_logp": -4.0}):
                beams = d.decode_beams(lg, **kw)
                beam_cases.append({"logits": f"logits_{k}", "decoder": name, "kwargs": kw,
                                   "beams": [[b[0], b[3], b[4]] for b in beams]})
    np.savez_compressed(os.path.join(HERE, "beam_toy_logits.npz"), **logits_store)
    json.dump({"labels": labels, "cases": beam_cases}, open(os.path.join(HERE, "beam_toy.json"), "w"),
              ensure_ascii=False, indent=0)

    known = [
        # (ref, hyp, chars SDIH, words SDIH) -- first rows are SURVEY 8c's hand-checked answers
        ("ab", "ba"), ("abc", "bcd"), ("hej med dig", ""), ("hej med dig", "hej  med   dig"),
        ("hej med dig", "hej med dig"), ("a b c d", "a x c"), ("kat", "skat"), ("  hej  ", "hej"),
        ("en to tre fire", "en tre to fire fem"), ("æble ø å", "aeble ø"), ("a\tb c", "a b c"),
    ]
    rng = np.random.default_rng(3)
    for _ in range(60):
        a = "".join(rng.choice(list("abcde "), size=int(rng.integers(1, 40)))).strip() or "a"
        b = "".join(rng.choice(list("abcde "), size=int(rng.integers(0, 40))))
        known.append((a, b))
    out = [{"ref": r, "hyp": h, "chars": list(edit.char_counts(r, h)), "words": list(edit.word_counts(r, h))}
           for r, h in known]
    json.dump(out, open(os.path.join(HERE, "edit_known.json"), "w"), ensure_ascii=False, indent=0)


if __name__ == "__main__":
    make_greedy()
    make_oracle_anchors()
    print("golden fixtures written to", HERE)
