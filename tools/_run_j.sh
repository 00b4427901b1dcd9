nvidia-smi topo -m > gpurun_out/r2_j_topo.txt 2>&1; nproc >> gpurun_out/r2_j_topo.txt; lscpu | grep -i "numa\|model name\|socket" >> gpurun_out/r2_j_topo.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_j_bench8.json 2> gpurun_out/r2_j_bench8.err
tail -3 gpurun_out/r2_j_bench8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_j_bench8.json").read().strip().splitlines()[-1])
print("N=8 value",round(d["value"]), "e2e",round(d["e2e"]["value"]), "list",round(d["e2e"]["list_input"]["value"]), "kernel_ms", round(d["roofline"]["kernel_ms_per_launch"],2), d["e2e"]["phases_ms"])
PY
tail -8 gpurun_out/r2_j_topo.txt
