python bench.py --steps 10 --warmup 3 > gpurun_out/r2_e_bench.json 2> gpurun_out/r2_e_bench.err
tail -2 gpurun_out/r2_e_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_e_bench.json").read().strip().splitlines()[-1])
print("value",round(d["value"]), "e2e",round(d["e2e"]["value"]), "list",round(d["e2e"]["list_input"]["value"]), "kernel_ms", round(d["roofline"]["kernel_ms_per_launch"],2), d["e2e"]["phases_ms"], d["other_kernels"], d["cpu_baseline"]["value"])
PY
ncu --set full --clock-control none --import-source on -k regex:beam_search_kernel -s 2 -c 1 -o gpurun_out/r2_e_beam python tools/beam_perf.py --utts 8192 --iters 1 > gpurun_out/r2_e_ncu.log 2>&1
tail -3 gpurun_out/r2_e_ncu.log
