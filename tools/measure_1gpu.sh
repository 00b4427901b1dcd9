set -x
mkdir -p gpurun_out/final
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final/r2_bench_reference.json 2> gpurun_out/final/r2_bench_reference.err
python bench.py --steps 20 --warmup 3 > gpurun_out/final/r2_bench_ours.json 2> gpurun_out/final/r2_bench_ours.err
python bench.py --steps 5 --warmup 3 --kind flat --utts 1776 --no-cpu-baseline > gpurun_out/final/r2_bench_flat.json 2> gpurun_out/final/r2_bench_flat.err
for c in 3 4 5; do python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/final/r2_bench_config$c.json 2> gpurun_out/final/r2_bench_config$c.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final/r2_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:beam_search_kernel -s 4 -c 1 -o gpurun_out/final/r2_beam_lean python tools/beam_perf.py --utts 8192 --iters 1 > gpurun_out/final/ncu_beam.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ctc_greedy_fused -s 3 -c 1 -o gpurun_out/final/r2_greedy python - > gpurun_out/final/ncu_greedy.log 2>&1 <<'PY'
import sys, os, tempfile, torch
sys.path.insert(0, os.getcwd())
import synth
from coral_b200.greedy import greedy_decode_device
wl = synth.build_workload(os.path.join(tempfile.gettempdir(), "coral_b200_cache"), 8192, order=5, name="eval0")
d = torch.from_numpy(wl.logits).cuda(); l = torch.from_numpy(wl.lengths).cuda()
for _ in range(5): greedy_decode_device(d, l, blank_id=45)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:edit_bitpar -s 2 -c 1 -o gpurun_out/final/r2_edit python tools/validation_perf.py 200000 > gpurun_out/final/ncu_edit.log 2>&1
python tools/validation_perf.py 200000 > gpurun_out/final/r2_validation_perf.json 2>&1
python tools/greedy_perf.py > gpurun_out/final/r2_greedy_roofline.jsonl 2>&1
ls -la gpurun_out/final
