python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "greedy or golden or compute_error or kenlm" 2>&1 | tail -3 > gpurun_out/r2_i_tests.log
python tools/greedy_perf.py > gpurun_out/r2_i_greedy.jsonl 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_i_bench.json 2> gpurun_out/r2_i_bench.err
cat gpurun_out/r2_i_tests.log; cat gpurun_out/r2_i_greedy.jsonl; tail -2 gpurun_out/r2_i_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_i_bench.json").read().strip().splitlines()[-1])
print("value",round(d["value"]), "e2e",round(d["e2e"]["value"]), "list",round(d["e2e"]["list_input"]["value"]), "kernel_ms", round(d["roofline"]["kernel_ms_per_launch"],2), d["e2e"]["phases_ms"], d["other_kernels"])
PY
