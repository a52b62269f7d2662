#!/bin/bash
# compute-sanitizer over a small parity subset (memcheck, then racecheck on shared memory).  Slow: keep the subset small.
mkdir -p gpurun_out
SUB='corners_bit_exact and (kat or ties or one_bright) and not 4000'
timeout -k 10 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "($SUB) or klt_kernel_modes or pair_frontend_batch or ransac_counts or multitracker_equals" > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/memcheck.log | tail -2 | tr '\n' ' ')" | tee -a gpurun_out/summary.txt
timeout -k 10 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "(corners_bit_exact and kat and 2200) or pair_frontend_batch_with or (ransac_counts and 300)" > gpurun_out/racecheck.log 2>&1
echo "racecheck rc=$? $(grep -E 'RACECHECK SUMMARY|passed|failed' gpurun_out/racecheck.log | tail -2 | tr '\n' ' ')" | tee -a gpurun_out/summary.txt
