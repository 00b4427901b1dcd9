"""BASELINE config 4: per-sample CER filter over 200k synthetic (reference, hypothesis) pairs.
Scoring only (hypotheses are 7 % corruptions); reports pairs/s end to end through the public API,
the edit kernel's own time, GCUPS and algorithmic GB/s."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import synth
from coral_b200 import metrics
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

# ---- config 4, end-to-end leg: greedy decode of peaky logits from host memory + per-sample scores
if os.environ.get("CORAL_VALIDATION_GREEDY", "1") != "0":
    import tempfile
    from coral_b200.greedy import CTCVocabulary, greedy_decode
    vocab_tokens = [c for c in synth.CORAL_LABELS]
    cache = os.path.join(tempfile.gettempdir(), "coral_b200_cache")
    wl = synth.build_workload(cache, 8192, order=5, name="eval0")
    labels = ["|" if c == " " else ("<pad>" if c == "" else ("<unk>" if c == "⁇" else c)) for c in wl.labels]
    vocab = CTCVocabulary(labels, labels.index("<pad>"))
    h_logits = torch.from_numpy(wl.logits).pin_memory()
    for _ in range(2):
        hy = greedy_decode(h_logits, vocab, lengths=wl.lengths)
        validation_scores(hy, wl.references)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    hy = greedy_decode(h_logits, vocab, lengths=wl.lengths)
    t_dec = time.perf_counter() - t0
    vs2 = validation_scores(hy, wl.references, max_cer=0.6)
    t_all = time.perf_counter() - t0
    print(json.dumps({"config": 4, "leg": "greedy decode from pinned host logits + per-sample CER/WER + keep mask",
                      "utterances": 8192, "decode_s": round(t_dec, 4), "total_s": round(t_all, 4),
                      "utt_per_s": round(8192 / t_all), "h2d_MB": round(h_logits.numel() * 4 / 1e6),
                      "cer": vs2.cer, "kept_fraction": float(vs2.keep.mean())}))
