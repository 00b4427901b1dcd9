ncu --set full --clock-control none --import-source on -k regex:beam_search_kernel -s 2 -c 1 -o gpurun_out/r2_m_flat python tools/beam_perf.py --utts 512 --kind flat --iters 1 > gpurun_out/r2_m_ncu.log 2>&1
tail -2 gpurun_out/r2_m_ncu.log
