python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "greedy or golden or compute_error" 2>&1 | tail -6 > gpurun_out/r2_g_tests.log
python tools/greedy_perf.py > gpurun_out/r2_g_greedy.jsonl 2>&1
python - > gpurun_out/r2_g_ragged.log 2>&1 <<'PY'
import sys, os, tempfile, numpy as np, torch
sys.path.insert(0, os.getcwd())
import synth
from coral_b200.greedy import greedy_decode_device
wl = synth.build_workload(os.path.join(tempfile.gettempdir(), "coral_b200_cache"), 8192, order=5, name="eval0")
d = torch.from_numpy(wl.logits).cuda(); l = torch.from_numpy(wl.lengths).cuda()
for _ in range(3): greedy_decode_device(d, l, blank_id=45)
ts=[]
for _ in range(9):
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record(); greedy_decode_device(d, l, blank_id=45); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
fr=int(wl.lengths.sum()); by=fr*46*4+fr*4
ms=float(np.median(ts)); print("ragged 8192: %.4f ms  %.0f GB/s  frac %.3f" % (ms, by/ms/1e6, by/ms/1e6/6555.5), ts)
PY
cat gpurun_out/r2_g_tests.log; cat gpurun_out/r2_g_greedy.jsonl; cat gpurun_out/r2_g_ragged.log
