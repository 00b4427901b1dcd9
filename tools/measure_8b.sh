mkdir -p gpurun_out/final
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --config 4 --steps 5 --warmup 3 > gpurun_out/final/r2_bench_config4_8gpu.json 2> gpurun_out/final/r2_bench_config4_8gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/final/r2_bench_config4_8gpu.json").read().strip().splitlines()[-1]); print("cfg4 N=8 value",round(d["value"]),"e2e",round(d["e2e"]["value"]), d["quality"])
PY
