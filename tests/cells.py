"""Parity cells for the named BASELINE configs (test infrastructure).

A *cell* = (seeded synthetic workload, decoder arguments). ``cell_inputs`` rebuilds the inputs
deterministically (numpy generators, the in-repo ARPA estimator); ``oracle_beams`` returns the
oracle's full beam lists for the cell -- from ``tests/golden/cells/<name>.json.gz`` when the
committed file was generated from exactly these inputs (SHA-256 fingerprint of the logits, the
ARPA file and the arguments), otherwise by running ``oracle.beam`` on a fork pool.

The cached files exist only because a few cells cost minutes of CPython per run (flat
posteriors, ``token_min_logp=-20``); ``tests/golden/make_cells.py`` is the generating script and
``tests/test_oracle_golden.py::test_cell_cache_matches_live_oracle`` re-derives a sample of every
cached cell on the CPU so a stale file cannot pass silently.

Cells (SURVEY.md section 8d / BASELINE.json configs):
  c2flat            config 2 stress variant: flat logits, read-aloud durations (T up to 499), beam 100
  c3_tml{3,5,7,10,20}  config 3: conversation durations (T up to 1499), beam 200, token_min_logp sweep
  c3flat_tml{3,5}   config 3 on flat logits
  c5_o{3,4,5,6}_b{16,32,64,128,256,512}  config 5: LM order x beam width, 50k-word LM
"""

from __future__ import annotations

import gzip
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CACHE = os.environ.get("CORAL_B200_CACHE", os.path.join(tempfile.gettempdir(), "coral_b200_cache"))
CELL_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cells")
N_UTTS = 16


def _cells():
    c = {}
    c["c2flat"] = dict(wl=dict(n=96, order=5, kind="flat", shape="read_aloud", name="cell2flat"),
                       kw=dict(beam_width=100, beam_prune_logp=-10.0, token_min_logp=-5.0))
    for t in (3, 5, 7, 10, 20):
        c[f"c3_tml{t}"] = dict(wl=dict(n=96, order=5, kind="peaky", shape="conversation", name="cell3"),
                               kw=dict(beam_width=200, beam_prune_logp=-10.0, token_min_logp=-float(t)))
    for t in (3, 5):
        c[f"c3flat_tml{t}"] = dict(wl=dict(n=96, order=5, kind="flat", shape="conversation", name="cell3flat"),
                                   kw=dict(beam_width=200, beam_prune_logp=-10.0, token_min_logp=-float(t)))
    for o in (3, 4, 5, 6):
        for b in (16, 32, 64, 128, 256, 512):
            c[f"c5_o{o}_b{b}"] = dict(wl=dict(n=96, order=o, kind="peaky", shape="read_aloud", name="cell5"),
                                      kw=dict(beam_width=b, beam_prune_logp=-10.0, token_min_logp=-5.0))
    return c


CELLS = _cells()
# cells whose oracle pass is too slow to repeat on every test run: their beams are committed
CACHED = ("c2flat", "c3_tml20", "c3flat_tml3", "c3flat_tml5")

_WL = {}


def cell_inputs(name: str):
    """(labels, arpa_path, [logits [T_i, 46] float32] * N_UTTS, kwargs). The utterances are the
    longest one of the workload plus evenly spaced length quantiles, so T_max of the config
    (499 / 1499 frames) is always covered."""
    import synth

    spec = CELLS[name]
    key = json.dumps(spec["wl"], sort_keys=True)
    if key not in _WL:
        w = spec["wl"]
        _WL[key] = synth.build_workload(CACHE, w["n"], order=w["order"], kind=w["kind"], shape=w["shape"],
                                        name=w["name"])
    wl = _WL[key]
    order = np.argsort(-wl.lengths, kind="stable")
    pick = order[np.linspace(0, len(order) - 1, N_UTTS).astype(int)]
    logits = [np.ascontiguousarray(wl.logits[u, : wl.lengths[u]]) for u in pick]
    return wl.labels, wl.arpa_path, logits, dict(spec["kw"])


def fingerprint(name: str) -> str:
    labels, arpa, logits, kw = cell_inputs(name)
    h = hashlib.sha256()
    h.update(json.dumps([labels, kw], sort_keys=True).encode())
    with open(arpa, "rb") as f:
        h.update(hashlib.sha256(f.read()).digest())
    for x in logits:
        h.update(np.int64(x.shape[0]).tobytes())
        h.update(x.tobytes())
    return h.hexdigest()


# ----------------------------------------------------------------------------- live oracle
_DEC = {}


def _init(labels, arpa):
    from oracle.beam import build_ctcdecoder

    _DEC["d"] = build_ctcdecoder(labels, arpa)


def _one(args):
    x, kw = args
    beams = _DEC["d"].decode_beams(x, **kw)
    return [[t, [[w, [int(a), int(b)]] for w, (a, b) in fr], float(ls), float(cs)] for t, _, fr, ls, cs in beams]


def live_oracle(name: str, which=None, n_procs: int | None = None):
    """Oracle beams of the cell's utterances ``which`` (default all): list per utterance of
    ``[text, [[word, [start, end]], ...], logit_score, lm_score]``."""
    import multiprocessing as mp

    labels, arpa, logits, kw = cell_inputs(name)
    idx = list(range(len(logits))) if which is None else list(which)
    items = [(logits[i], kw) for i in idx]
    n_procs = n_procs or min(len(items), os.cpu_count() or 1)
    if n_procs <= 1:
        _init(labels, arpa)
        return [_one(it) for it in items]
    with mp.get_context("fork").Pool(n_procs, initializer=_init, initargs=(labels, arpa)) as pool:
        return pool.map(_one, items, chunksize=1)


def cache_path(name: str) -> str:
    return os.path.join(CELL_DIR, name + ".json.gz")


def load_cache(name: str):
    p = cache_path(name)
    if not os.path.exists(p):
        return None
    with gzip.open(p, "rt", encoding="utf-8") as f:
        return json.load(f)


def _tmp_path(name: str) -> str:
    return os.path.join(CACHE, "cells", f"{name}_{fingerprint(name)[:16]}.json.gz")


def precompute(names) -> None:
    """Run the oracle for ``names`` (fork pool over utterances) and park the beams in the scratch
    cache. Called in a fresh subprocess by ``ensure`` so that no CUDA-initialised process forks."""
    os.makedirs(os.path.join(CACHE, "cells"), exist_ok=True)
    for name in names:
        p = _tmp_path(name)
        if os.path.exists(p):
            continue
        data = {"fingerprint": fingerprint(name), "beams": live_oracle(name)}
        tmp = p + f".tmp{os.getpid()}"
        with gzip.open(tmp, "wt", encoding="utf-8", compresslevel=1) as f:
            json.dump(data, f, ensure_ascii=False)
        os.replace(tmp, p)


def ensure(names) -> None:
    """Make sure oracle beams exist for every cell in ``names`` (committed file with a matching
    fingerprint, or a scratch file computed now in a subprocess)."""
    import subprocess

    todo = []
    for name in names:
        data = load_cache(name) if name in CACHED else None
        if data is not None and data.get("fingerprint") == fingerprint(name):
            continue
        if not os.path.exists(_tmp_path(name)):
            todo.append(name)
    if todo:
        subprocess.run([sys.executable, os.path.abspath(__file__), *todo], check=True)


def oracle_beams(name: str):
    """Oracle beam lists of the cell as 5-tuples ``(text, None, [(word, (start, end))], logit, lm)``
    (the shape ``conftest.beams_equal`` takes) and where they came from ("golden" / "live")."""
    data = load_cache(name) if name in CACHED else None
    src = "golden"
    if data is None or data.get("fingerprint") != fingerprint(name):
        ensure([name])
        with gzip.open(_tmp_path(name), "rt", encoding="utf-8") as f:
            data = json.load(f)
        src = "live"
    out = [[(t, None, [(w, (a, b)) for w, (a, b) in fr], ls, cs) for t, fr, ls, cs in utt] for utt in data["beams"]]
    return out, src


if __name__ == "__main__":
    precompute(sys.argv[1:])
