python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_c_tests.log
CORAL_PHASES=1 python tools/beam_perf.py --utts 8192 --iters 7 > gpurun_out/r2_c_perf.log 2>&1
CORAL_B200_LIB=coral_b200/lib/ab/libcoral_b200_noblank.so python tools/beam_perf.py --utts 8192 --iters 7 >> gpurun_out/r2_c_perf.log 2>&1
python tools/beam_perf.py --utts 1776 --kind flat --iters 3 >> gpurun_out/r2_c_perf.log 2>&1
python tools/beam_perf.py --utts 2048 --beam 256 --iters 3 >> gpurun_out/r2_c_perf.log 2>&1
python tools/beam_perf.py --utts 2048 --beam 512 --iters 3 >> gpurun_out/r2_c_perf.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_c_bench.json 2> gpurun_out/r2_c_bench.err
cat gpurun_out/r2_c_tests.log; grep -v "^$" gpurun_out/r2_c_perf.log; tail -2 gpurun_out/r2_c_bench.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_c_bench.json").read().strip().splitlines()[-1])
print("value",round(d["value"]), "e2e",round(d["e2e"]["value"]), "list",round(d["e2e"]["list_input"]["value"]), "kernel_ms", round(d["roofline"]["kernel_ms_per_launch"],2), d["e2e"]["phases_ms"], d["other_kernels"])
PY
