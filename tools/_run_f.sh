python -m pytest tests/test_gpu_parity.py tests/test_independent_pins.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_f_tests.log
python tools/beam_perf.py --utts 8192 --iters 7 > gpurun_out/r2_f_perf.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_f_bench.json 2> gpurun_out/r2_f_bench.err
cat gpurun_out/r2_f_tests.log; grep utts gpurun_out/r2_f_perf.log; tail -2 gpurun_out/r2_f_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_f_bench.json").read().strip().splitlines()[-1])
print("value",round(d["value"]), "e2e",round(d["e2e"]["value"]), "list",round(d["e2e"]["list_input"]["value"]), "kernel_ms", round(d["roofline"]["kernel_ms_per_launch"],2), d["e2e"]["phases_ms"])
PY
