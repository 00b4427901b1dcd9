"""Small decode used under compute-sanitizer (racecheck / memcheck)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import synth
from coral_b200.decoder import build_ctcdecoder
cache = os.path.join(tempfile.gettempdir(), "coral_b200_cache")
w = synth.build_workload(cache, 4, order=4, n_words=2000, n_sent=5000, name="t")
dec = build_ctcdecoder(w.labels, w.arpa_path)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 60
lg = [w.logits[u, : min(T, w.lengths[u])] for u in range(4)]
out = dec.decode_beams_batch(None, lg)
print([len(b) for b in out], out[0][0][0][:50])
