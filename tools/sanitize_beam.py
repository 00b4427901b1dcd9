"""Workload for compute-sanitizer (memcheck / racecheck) over the beam-search kernels: classify +
lean + heavy (flat logits), the word-frame instantiations, prune_history, beams 16 / 100 / 256 / 512,
pinned host logits read in place, a list of arrays through the staging buffer, decode_batches.
`compute-sanitizer --tool memcheck python tools/sanitize_beam.py [frames]`."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import synth
from coral_b200.decoder import build_ctcdecoder

cache = os.path.join(tempfile.gettempdir(), "coral_b200_cache")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 48
n = 12
w = synth.build_workload(cache, n, order=4, n_words=2000, n_sent=5000, name="t")
f = synth.build_workload(cache, n, order=4, n_words=2000, n_sent=5000, name="t", kind="flat")
dec = build_ctcdecoder(w.labels, w.arpa_path)
cut = lambda wl: [wl.logits[u, : min(T, wl.lengths[u])] for u in range(n)]
outs = []
for wl in (w, f):
    for beam in (16, 100, 256, 512):
        outs.append(dec.decode_beams_batch(None, cut(wl), beam_width=beam, n_best=3))      # word-frame kernels
        outs.append(dec.decode_batch(None, cut(wl), beam_width=beam))                      # text kernels, list input
    outs.append(dec.decode_beams_batch(None, cut(wl), prune_history=True, n_best=2))
    pinned = torch.from_numpy(np.ascontiguousarray(wl.logits[:, :T])).pin_memory()
    lens = np.minimum(wl.lengths, T)
    a = dec.decode_batch(None, pinned, lengths=lens)
    b = dec.decode_batch(None, pinned.cuda(), lengths=lens)
    assert list(a) == list(b), "pinned host logits and device logits decode differently"
    got = [list(x) for x in dec.decode_batches([(pinned, lens), cut(wl), (pinned.cuda(), lens)])]
    assert got[0] == got[1] == got[2] == list(a), "decode_batches differs"
torch.cuda.synchronize()
print("sanitizer workload ok", len(outs))
