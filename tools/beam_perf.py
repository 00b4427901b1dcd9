"""Kernel-level timing of the beam search on a fixed synthetic workload (tuning helper)."""
import argparse, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import synth
from coral_b200.decoder import build_ctcdecoder

ap = argparse.ArgumentParser()
ap.add_argument("--utts", type=int, default=2048)
ap.add_argument("--beam", type=int, default=100)
ap.add_argument("--order", type=int, default=5)
ap.add_argument("--kind", default="peaky")
ap.add_argument("--shape", default="read_aloud")
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--mode", type=int, default=0)
a = ap.parse_args()
cache = os.path.join(tempfile.gettempdir(), "coral_b200_cache")
wl = synth.build_workload(cache, a.utts, order=a.order, kind=a.kind, shape=a.shape, name="eval0")
dec = build_ctcdecoder(wl.labels, wl.arpa_path)
dev = torch.device("cuda", 0)
d_logits = torch.from_numpy(wl.logits).to(dev); d_len = torch.from_numpy(wl.lengths).to(dev)
d_order = torch.argsort(d_len, descending=True).to(torch.int32)
for _ in range(2):
    dec.decode_launch(d_logits, d_len, d_order, beam_width=a.beam, input_mode=a.mode)
torch.cuda.synchronize()
ts = []
for _ in range(a.iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dec.decode_launch(d_logits, d_len, d_order, beam_width=a.beam, input_mode=a.mode, events=(e0, e1))
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
frames = int(wl.lengths.sum())
ms = float(np.median(ts))
print(f"NT={os.environ.get('CORAL_BEAM_NT','default')} utts={a.utts} beam={a.beam} kind={a.kind}: {ms:.2f} ms  "
      f"{a.utts/ms*1e3:.0f} utt/s  {ms*1e3/frames*1e3:.1f} ns/frame-amortised  (min {min(ts):.2f})")
if os.environ.get("CORAL_PHASES"):
    out = dec.decode_padded(d_logits, d_len, beam_width=a.beam, n_best=1, input_mode=a.mode, collect_stats=True)
    st = out.stats.astype(np.float64)
    fr = st[3]
    names = ["hash", "expand", "ovf-select", "bucket", "scatter", "rank+commit", "grow", "stage"]
    tot = st[8:16].sum()
    print("per-frame cycles (thread 0): " + ", ".join(f"{n}={st[8+i]/fr:.0f}" for i, n in enumerate(names)) + f"  total={tot/fr:.0f}")
    ops = ["-", "lm_word_score", "-", "lex_find", "-"]
    print("op latency (cycles/call, calls/frame): " + ", ".join(
        f"{n}={st[16+i]/max(st[24+i],1):.0f}x{st[24+i]/fr:.2f}" for i, n in enumerate(ops)))
    n0 = max(st[24], 1)
    print(f"letter item (cycles): lookup+child-hash {st[16]/n0:.0f}, merge {st[18]/n0:.0f}, lexicon+penalty {st[20]/n0:.0f}, emit {st[23]/n0:.0f}  x{st[24]/fr:.1f}/frame")
    print(f"expand cycles: frames with LM scoring {st[21]/max(st[29],1):.0f} x{st[29]/fr:.2f} of frames, without {st[22]/max(st[30],1):.0f} x{st[30]/fr:.2f}")
    print(f"per frame: ext={st[0]/fr:.1f} lm_scorings={st[1]/fr:.2f} ngram_probes={st[2]/fr:.2f} lex_probes={st[4]/fr:.2f} nodes={st[5]/fr:.2f} radix_select_frames/utt={st[7]/a.utts:.2f}")
if os.environ.get("CORAL_FRAMES"):
    ts = []
    for _ in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dec.decode_launch(d_logits, d_len, d_order, beam_width=a.beam, input_mode=a.mode, events=(e0, e1), word_frames=True)
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"with word frames: {float(np.median(ts)):.2f} ms (min {min(ts):.2f})")
