#!/usr/bin/env python
"""Writes REFERENCE-generated goldens wherever the real packages import.

    python tools/pin_oracle.py [--out tests/golden] [--bin path/to/model.bin ...]

pyctcdecode 0.5.0 / kenlm / jiwer 4.0.0 (-> rapidfuzz) are the reference's arithmetic for this
path (R:uv.lock:2357-2358, :1275-1278, :1204-1205) and are not installable in the build image.
On a machine that has them this script

1. runs the real ``jiwer.process_characters`` / ``process_words`` on a seeded set of pairs
   (incl. > 64-symbol strings, empty hypotheses, an empty reference) -> ``ref_edit.json``;
2. runs the real ``pyctcdecode`` decoder (with the real ``kenlm``) on the seeded synthetic
   utterances of tests/conftest.py and on flat logits -> ``ref_beam.json`` (all beams: text,
   word frames, logit and LM scores) and per-word ``kenlm`` scores -> ``ref_lm.json``;
3. for every KenLM binary given with ``--bin`` (or found under ``--search``): loads it with the
   real ``kenlm.Model`` and with this repo's reader (tests/hostsim, CPU) and compares the
   per-word scores of seeded sentences -> ``ref_kenlm_bin.json`` (sha256 + scores).

tests/test_reference_goldens.py consumes these files when they exist (CPU: oracle == file;
GPU: CUDA path == file). The files carry the package versions that produced them. Without the
packages the script says so and exits 3: nothing is ever written from the oracle here.
"""

from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def edit_pairs(seed=11):
    rng = np.random.default_rng(seed)
    letters = np.array(list("abcdeæøå") + [" "] * 3 + ["\t"])
    pairs = [("ab", "ba"), ("abc", "bcd"), ("hej med dig", "hej  med   dig"), ("a b c d", "a x c"), ("abc", ""),
             ("short one here", "shoe order one"), ("ab", "bca"), ("ab", "ca")]
    for hi in (12, 40, 90, 200, 400):
        for _ in range(120):
            r = "".join(letters[rng.integers(0, len(letters), size=int(rng.integers(1, hi)))]).strip() or "a"
            h = list(r)
            for _ in range(int(rng.integers(0, max(2, len(r) // 4)))):
                k, pos = int(rng.integers(0, 3)), int(rng.integers(0, len(h) + 1))
                if k == 0 and h:
                    h[min(pos, len(h) - 1)] = str(letters[rng.integers(0, len(letters))])
                elif k == 1 and h:
                    del h[min(pos, len(h) - 1)]
                else:
                    h.insert(pos, str(letters[rng.integers(0, len(letters))]))
            pairs.append((r, "".join(h)))
    return pairs


def versions():
    from importlib import metadata

    out = {}
    for name in ("pyctcdecode", "kenlm", "jiwer", "rapidfuzz", "pygtrie", "numpy"):
        try:
            out[name] = metadata.version(name)
        except Exception:
            out[name] = None
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--bin", action="append", default=[], help="a KenLM binary produced by the real build_binary")
    ap.add_argument("--search", default=None, help="directory to scan for *.bin KenLM files")
    args = ap.parse_args()

    from oracle import selfcheck

    if not selfcheck.real_packages_available():
        print("pyctcdecode / kenlm / jiwer are not importable here: nothing written "
              "(goldens from this script must come from the real packages)")
        return 3
    import jiwer
    import kenlm
    from pyctcdecode import build_ctcdecoder as real_build

    import synth

    meta = {"generator": "tools/pin_oracle.py", "versions": versions()}
    os.makedirs(args.out, exist_ok=True)

    # ---- 1. jiwer / rapidfuzz
    rows = []
    for ref, hyp in edit_pairs():
        c = jiwer.process_characters(reference=ref, hypothesis=hyp)
        w = jiwer.process_words(reference=ref, hypothesis=hyp)
        rows.append({"ref": ref, "hyp": hyp, "chars": [c.substitutions, c.deletions, c.insertions, c.hits],
                     "words": [w.substitutions, w.deletions, w.insertions, w.hits]})
    empty = {}
    for kind, fn in (("chars", jiwer.process_characters), ("words", jiwer.process_words)):
        try:
            m = fn(reference="", hypothesis="abc def")
            empty[kind] = [m.substitutions, m.deletions, m.insertions, m.hits]
        except Exception as e:  # jiwer < 3.1 raises ValueError here
            empty[kind] = {"raises": type(e).__name__}
    json.dump({"meta": meta, "rows": rows, "empty_reference": empty},
              open(os.path.join(args.out, "ref_edit.json"), "w"), ensure_ascii=False, indent=0)

    # ---- 2. pyctcdecode + kenlm on the seeded synthetic utterances of tests/conftest.py
    cache = os.environ.get("CORAL_B200_CACHE", os.path.join(tempfile.gettempdir(), "coral_b200_cache"))
    words, model, arpa = synth.build_lm(cache, order=4, n_words=2000, n_sent=5000)
    wl = synth.build_workload(cache, 12, order=4, n_words=2000, n_sent=5000, name="t")
    dec = real_build(list(synth.CORAL_LABELS), kenlm_model_path=arpa)
    rng = np.random.default_rng(5)
    cases = []
    inputs = [("t%d" % u, wl.logits[u, : wl.lengths[u]], {}) for u in range(12)]
    inputs += [("flat%d" % k, synth.flat_logits(40 + 20 * k, rng), {"beam_width": bw})
               for k, bw in enumerate((16, 100, 200))]
    inputs += [("t0_prune", wl.logits[0, : wl.lengths[0]], {"prune_history": True}),
               ("t1_params", wl.logits[1, : wl.lengths[1]], {"beam_width": 25, "beam_prune_logp": -5.0, "token_min_logp": -3.0})]
    for name, lg, kw in inputs:
        beams = dec.decode_beams(np.asarray(lg), **kw)
        cases.append({"name": name, "kwargs": kw, "sha256": hashlib.sha256(np.ascontiguousarray(lg).tobytes()).hexdigest(),
                      "beams": [[b[0], [[w, [int(f[0]), int(f[1])]] for w, f in b[2]], float(b[3]), float(b[4])] for b in beams]})
    json.dump({"meta": meta, "workload": "tests/conftest.py small_workload (order 4, 2000 words, 5000 sentences, name 't') + synth.flat_logits(seed 5)",
               "cases": cases}, open(os.path.join(args.out, "ref_beam.json"), "w"), ensure_ascii=False)
    km = kenlm.Model(arpa)
    flat, lens = model.sample(200, "lmtest")
    sents = synth.sentences_to_text(flat, lens, words)
    lm_rows = []
    for k, s in enumerate(sents):
        ws = s.split(" ")
        if k % 3 == 0:
            ws[len(ws) // 2] = "zzzoov"
        s = " ".join(ws)
        lm_rows.append({"sentence": s, "full_scores": [[float(p), int(n), bool(o)] for p, n, o in km.full_scores(s, bos=True, eos=True)]})
    json.dump({"meta": meta, "arpa": "synth.build_lm(order=4, n_words=2000, n_sent=5000)", "rows": lm_rows},
              open(os.path.join(args.out, "ref_lm.json"), "w"), ensure_ascii=False)

    # ---- 3. real KenLM binaries
    bins = list(args.bin)
    if args.search:
        for d, _, fs in os.walk(args.search):
            bins += [os.path.join(d, f) for f in fs if f.endswith(".bin")]
    bin_rows = []
    for path in bins:
        try:
            real = kenlm.Model(path)
        except Exception as e:
            print("skipping", path, e)
            continue
        from hostsim_lib import HostSim

        ours = HostSim([""], path, unigrams=None)  # CPU build of this repo's reader + scorer
        rows = []
        ok = True
        for s in sents[:100]:
            ws = s.split(" ")
            ref = [float(np.float32(p)) for p, _, _ in real.full_scores(s, bos=True, eos=True)]
            got, _ = ours.score_sentence(ws)
            ok = ok and [float(x) for x in got] == ref
            rows.append({"sentence": s, "log10": ref})
        bin_rows.append({"path": os.path.basename(path), "sha256": hashlib.sha256(open(path, "rb").read()).hexdigest(),
                         "order": real.order, "reader_matches_kenlm": ok, "rows": rows})
        print(path, "reader == kenlm.Model:", ok)
    if bin_rows:
        json.dump({"meta": meta, "binaries": bin_rows}, open(os.path.join(args.out, "ref_kenlm_bin.json"), "w"),
                  ensure_ascii=False)
    print("wrote ref_edit.json, ref_beam.json, ref_lm.json" + (", ref_kenlm_bin.json" if bin_rows else ""), "to", args.out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
