python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_d_tests_default.log
CORAL_B200_LIB=coral_b200/lib/ab/libcoral_b200_noblank.so python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_d_tests_noblank.log
python tools/beam_perf.py --utts 8192 --iters 7 > gpurun_out/r2_d_perf.log 2>&1
CORAL_B200_LIB=coral_b200/lib/ab/libcoral_b200_noblank.so python tools/beam_perf.py --utts 8192 --iters 7 >> gpurun_out/r2_d_perf.log 2>&1
python tools/beam_perf.py --utts 8192 --iters 7 >> gpurun_out/r2_d_perf.log 2>&1
CORAL_B200_LIB=coral_b200/lib/ab/libcoral_b200_noblank.so python tools/beam_perf.py --utts 8192 --iters 7 >> gpurun_out/r2_d_perf.log 2>&1
echo DEFAULT; cat gpurun_out/r2_d_tests_default.log; echo NOBLANK; cat gpurun_out/r2_d_tests_noblank.log; grep utts gpurun_out/r2_d_perf.log
