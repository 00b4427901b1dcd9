"""Workload for compute-sanitizer (memcheck / racecheck) over both edit kernels and every NW
instantiation: `compute-sanitizer --tool memcheck python tools/sanitize_edit.py`; checks every 9th
pair against the oracle as it goes."""

import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
from coral_b200 import metrics
from oracle import edit as oe
rng = np.random.default_rng(7)
def pairs(n, maxlen, alpha="abcdeæøå  \t"):
    refs, hyps = [], []
    for _ in range(n):
        r = "".join(rng.choice(list(alpha), size=int(rng.integers(0, maxlen))))
        if rng.random() < 0.6:
            h = list(r)
            for _ in range(int(rng.integers(0, 6))):
                k = int(rng.integers(0, len(h) + 1)); x = rng.random()
                if x < .3 and h: h.pop(min(k, len(h) - 1))
                elif x < .6: h.insert(k, rng.choice(list(alpha)))
                elif h: h[min(k, len(h) - 1)] = rng.choice(list(alpha))
            h = "".join(h)
        else:
            h = "".join(rng.choice(list(alpha), size=int(rng.integers(0, maxlen))))
        refs.append(r); hyps.append(h)
    return refs, hyps
for maxlen, n in ((60, 700), (125, 700), (190, 400), (250, 400), (400, 200)):
    refs, hyps = pairs(n, maxlen)
    for kind, f in (("chars", oe.char_counts), ("words", oe.word_counts)):
        got = metrics.edit_counts(hyps, refs, kind)
        for i in range(0, n, 9):
            assert tuple(got[i]) == f(refs[i], hyps[i]), (maxlen, kind, i)
torch.cuda.synchronize()
print("sanitizer workload ok")
