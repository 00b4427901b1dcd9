for lib in "" coral_b200/lib/ab/libcoral_b200_old.so ""; do
  echo "LIB=$lib"; if [ -n "$lib" ]; then export CORAL_B200_LIB=$lib; else unset CORAL_B200_LIB; fi; python tools/beam_perf.py --utts 8192 --iters 7 2>&1 | grep utts
done
unset CORAL_B200_LIB
python tools/beam_perf.py --utts 1776 --kind flat --iters 3 2>&1 | grep utts
python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -4
