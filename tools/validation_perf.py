"""BASELINE config 4: per-sample CER filter over 200k synthetic (reference, hypothesis) pairs.
Scoring only (hypotheses are 7 % corruptions); reports pairs/s end to end through the public API,
the edit kernel's own time, GCUPS and algorithmic GB/s."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from coral_b200 import synth, metrics
from coral_b200.textio import encode_utf32
from coral_b200.validation import validation_scores

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
rng = np.random.default_rng(4242)
words = synth.make_word_list(20_000)
idx = rng.integers(0, len(words), size=(n, 14))
lens = rng.integers(3, 15, size=n)
refs = [" ".join(words[j] for j in idx[i, : lens[i]]) for i in range(n)]
hyps = [synth.corrupt_text(r, rng, 0.07) for r in refs]
dev = torch.device("cuda", 0)
validation_scores(hyps[:1000], refs[:1000])
torch.cuda.synchronize()
t0 = time.perf_counter(); vs = validation_scores(hyps, refs, max_cer=0.6); t1 = time.perf_counter() - t0
# kernel-only
r_cps, r_off = encode_utf32(refs); h_cps, h_off = encode_utf32(hyps)
d = [torch.from_numpy(x.view(np.int32) if x.dtype == np.uint32 else x).to(dev) for x in (r_cps, r_off, h_cps, h_off)]
mx = int(max(np.diff(r_off).max(), np.diff(h_off).max()))
res = {}
for mode, name in ((1, "chars"), (2, "words")):
    for _ in range(2):
        metrics.edit_counts_device(d[0], d[1], d[2], d[3], n, mode, mx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); metrics.edit_counts_device(d[0], d[1], d[2], d[3], n, mode, mx); e1.record(); torch.cuda.synchronize()
    res[name] = e0.elapsed_time(e1)
cells = float((np.diff(r_off) * np.diff(h_off)).sum())
alg = float((np.diff(r_off) + np.diff(h_off)).sum() * 4 + 16 * n)
print(json.dumps({"config": 4, "pairs": n, "max_len": mx, "e2e_s": round(t1, 3), "pairs_per_s_e2e": round(n / t1),
                  "kernel_ms_chars": round(res["chars"], 3), "kernel_ms_words": round(res["words"], 3),
                  "GCUPS_chars_full_matrix": round(cells / res["chars"] / 1e6, 1),
                  "algorithmic_GB_per_s_chars": round(alg / res["chars"] / 1e6, 1),
                  "cer": vs.cer, "wer": vs.wer, "kept_fraction": float(vs.keep.mean())}))
