CORAL_PHASES=1 python tools/beam_perf.py --utts 1776 --kind flat --iters 2 > gpurun_out/r2_l_flat_phases.log 2>&1
grep -v "^Found" gpurun_out/r2_l_flat_phases.log
