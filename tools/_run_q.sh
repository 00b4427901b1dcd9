ncu --set full --clock-control none --import-source on -k regex:beam_search_kernel -s 5 -c 1 -o gpurun_out/r2_q_flat python tools/beam_perf.py --utts 1776 --kind flat --iters 1 > gpurun_out/r2_q_ncu.log 2>&1
tail -2 gpurun_out/r2_q_ncu.log
