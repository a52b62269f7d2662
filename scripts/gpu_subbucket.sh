#!/bin/bash
# the sub-bucket walk forced by a small gather and few buckets (scratch build): corner / front-end tests on that library, then timing A/B
mkdir -p gpurun_out; : > gpurun_out/summary.txt
SFMGPU_LIB=$PWD/scratch_libs/libsfmgpu_T2048_B5.so timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 600 -x -k "corners or pair_frontend or candidates or tracker" > gpurun_out/sub_tests.log 2>&1
echo "tests(T2048_B5) rc=$? $(tail -1 gpurun_out/sub_tests.log)" | tee -a gpurun_out/summary.txt
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 600 -x -k "corners or pair_frontend or candidates or tracker" > gpurun_out/sel_tests.log 2>&1
echo "tests(in-tree) rc=$? $(tail -1 gpurun_out/sel_tests.log)" | tee -a gpurun_out/summary.txt
bash scripts/ab_select.sh
