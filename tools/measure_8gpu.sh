mkdir -p gpurun_out/final
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/final/r2_bench_ours_8gpu.json 2> gpurun_out/final/r2_bench_ours_8gpu.err
$TR --master-port 29512 bench.py --gpus 8 --config 4 --steps 5 --warmup 3 > gpurun_out/final/r2_bench_config4_8gpu.json 2> gpurun_out/final/r2_bench_config4_8gpu.err
$TR --master-port 29513 bench.py --gpus 8 --config 5 --steps 3 --warmup 3 > gpurun_out/final/r2_bench_config5_8gpu.json 2> gpurun_out/final/r2_bench_config5_8gpu.err
tail -c 600 gpurun_out/final/r2_bench_ours_8gpu.json; tail -2 gpurun_out/final/*8gpu.err
