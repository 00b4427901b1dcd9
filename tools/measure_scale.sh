mkdir -p gpurun_out/final
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/final/r2_bench_ours_${N}gpu.json 2> gpurun_out/final/r2_bench_ours_${N}gpu.err
$TR --master-port 29522 bench.py --gpus $N --config 4 --steps 5 --warmup 3 > gpurun_out/final/r2_bench_config4_${N}gpu.json 2> gpurun_out/final/r2_bench_config4_${N}gpu.err
python - <<PY
import json
for f in ("gpurun_out/final/r2_bench_ours_${N}gpu.json","gpurun_out/final/r2_bench_config4_${N}gpu.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, "value",round(d["value"]),"e2e",round(d["e2e"]["value"]))
    except Exception as e: print(f,"ERR",e)
PY
