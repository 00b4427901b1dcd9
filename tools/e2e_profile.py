"""Where the end-to-end step (bench.py step_e2e) spends its time, host side included."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import synth
from coral_b200 import metrics
from coral_b200.decoder import build_ctcdecoder
cache = os.path.join(tempfile.gettempdir(), "coral_b200_cache")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
wl = synth.build_workload(cache, B, order=5, name="eval0")
dec = build_ctcdecoder(wl.labels, wl.arpa_path)
h_logits = torch.from_numpy(wl.logits).pin_memory(); h_len = torch.from_numpy(wl.lengths).pin_memory()
refs = wl.references
def sync(): torch.cuda.synchronize()
for rep in range(4):
    t = {}
    sync(); t0 = time.perf_counter()
    d = h_logits.cuda(non_blocking=True); sync(); t["(h2d alone)"] = time.perf_counter() - t0
    del d
    sync(); t0 = time.perf_counter()
    outs = dec.decode_padded(h_logits, h_len, n_best=1, to_host=False); t["decode_padded enqueue"] = time.perf_counter() - t0
    sync(); t["... until kernel done"] = time.perf_counter() - t0
    t1 = time.perf_counter(); hyps = dec.device_tokens_to_text(outs[3][:, 0, :], outs[4][:, 0]); t["device_tokens_to_text"] = time.perf_counter() - t1
    t1 = time.perf_counter(); st = outs[5].cpu().numpy(); t["status d2h"] = time.perf_counter() - t1
    t1 = time.perf_counter(); c = metrics.cer(hyps, refs); t["cer"] = time.perf_counter() - t1
    t1 = time.perf_counter(); w = metrics.wer(hyps, refs); t["wer"] = time.perf_counter() - t1
    tot = time.perf_counter() - t0
    print(f"rep{rep} total={tot*1e3:.1f} ms ({B/tot:.0f} utt/s): " + ", ".join(f"{k}={v*1e3:.2f}" for k, v in t.items()))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
hyps = dec.decode_batch(None, h_logits, lengths=h_len); c = metrics.cer(hyps, refs); w = metrics.wer(hyps, refs)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
