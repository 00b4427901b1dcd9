python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_b_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_b_bench_zc1.json 2> gpurun_out/r2_b_bench_zc1.err
CORAL_HOST_INPUT_MODE=2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b_bench_zc2.json 2> gpurun_out/r2_b_bench_zc2.err
CORAL_HOST_INPUT=dma python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b_bench_dma.json 2> gpurun_out/r2_b_bench_dma.err
tail -5 gpurun_out/r2_b_tests.log; tail -3 gpurun_out/r2_b_bench_zc1.err; python - <<'PY'
import json
for n in ("zc1","zc2","dma"):
    try:
        d=json.loads(open(f"gpurun_out/r2_b_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, "value",round(d["value"]), "e2e",round(d["e2e"]["value"]), "list",round(d["e2e"]["list_input"]["value"]), "kernel_ms", round(d["roofline"]["kernel_ms_per_launch"],2), d["e2e"]["phases_ms"])
    except Exception as e: print(n,"ERR",e)
PY
