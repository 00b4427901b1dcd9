"""HBM roofline of the greedy argmax + collapse kernels (the genuinely HBM-bound part of the path)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from coral_b200 import _lib
from coral_b200.greedy import greedy_decode_device

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
dev = torch.device("cuda", 0)
for (B, T) in ((64, 499), (8192, 499), (16384, 499)):
    V = 46
    logits = torch.randn((B, T, V), device=dev, dtype=torch.float32)
    lengths = torch.full((B,), T, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        greedy_decode_device(logits, lengths, blank_id=45)
    # back-to-back launches between two events: a host sync per launch would add the Python
    # launch overhead (~20 us) to a ~100 us kernel. The 1.5 GB / 0.75 GB inputs are far larger than L2.
    torch.cuda.synchronize()
    n_rep = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_rep):
        greedy_decode_device(logits, lengths, blank_id=45)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n_rep
    alg = B * T * V * 4 + B * T * 4  # every logit read once, at most one token written per frame
    print(json.dumps({"kernel": "ctc_greedy_fused_kernel", "B": B, "T": T, "ms": round(ms, 4),
                      "algorithmic_MB": round(alg / 1e6, 1), "GB_per_s": round(alg / ms / 1e6, 1),
                      "frac_of_measured_hbm": round(alg / ms / 1e6 / peak, 3), "utt_per_s": round(B / ms * 1e3)}))
