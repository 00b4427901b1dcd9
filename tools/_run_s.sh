python tools/beam_perf.py --utts 1776 --kind flat --iters 3 2>&1 | grep utts
CORAL_B200_LIB=coral_b200/lib/ab/libcoral_b200_div1.so python tools/beam_perf.py --utts 1776 --kind flat --iters 3 2>&1 | grep utts
python tools/beam_perf.py --utts 8192 --iters 5 2>&1 | grep utts
