"""Where the end-to-end step spends its time (host side included)."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from coral_b200 import synth, metrics
from coral_b200.decoder import build_ctcdecoder
cache = os.path.join(tempfile.gettempdir(), "coral_b200_cache")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
wl = synth.build_workload(cache, B, order=5, name="eval0")
dec = build_ctcdecoder(wl.labels, wl.arpa_path)
h_logits = torch.from_numpy(wl.logits).pin_memory(); h_len = torch.from_numpy(wl.lengths).pin_memory()
refs = wl.references
def sync(): torch.cuda.synchronize()
for rep in range(3):
    t = {}
    sync(); t0 = time.perf_counter()
    d = h_logits.cuda(non_blocking=True); sync(); t["h2d_only"] = time.perf_counter() - t0
    del d
    sync(); t0 = time.perf_counter()
    out = dec.decode_padded(h_logits, h_len, n_best=1); t["decode_padded(h2d+kernel+d2h)"] = time.perf_counter() - t0
    t0 = time.perf_counter(); hyps = dec.tokens_to_text(out.tokens[:, 0, :], out.lens[:, 0]); t["tokens_to_text"] = time.perf_counter() - t0
    t0 = time.perf_counter(); c = metrics.cer(hyps, refs); t["cer"] = time.perf_counter() - t0
    t0 = time.perf_counter(); w = metrics.wer(hyps, refs); t["wer"] = time.perf_counter() - t0
    tot = sum(v for k, v in t.items() if k != "h2d_only")
    print(f"rep{rep} total={tot*1e3:.1f} ms ({B/tot:.0f} utt/s): " + ", ".join(f"{k}={v*1e3:.1f}" for k, v in t.items()))
