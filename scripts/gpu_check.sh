#!/bin/bash
# Runs on the GPU box: smoke + the GPU parity tests, one pytest process per group so that a CUDA fault in one
# group cannot poison the others.  Logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout -k 10 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
for grp in generator pyramid candidates std_sort corners klt tracker ransac pair_frontend errors; do :; done
timeout -k 10 600 python -m pytest tests/test_gpu_shim.py -q -m gpu --timeout 300 -x > gpurun_out/test_shim.log 2>&1; echo "shim rc=$? $(tail -1 gpurun_out/test_shim.log)" | tee -a gpurun_out/summary.txt
for grp in generator pyramid candidates std_sort corners klt tracker ransac pair_frontend errors; do
  timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$grp" --timeout 300 -x -s > gpurun_out/test_$grp.log 2>&1
  echo "$grp rc=$? $(tail -1 gpurun_out/test_$grp.log)" | tee -a gpurun_out/summary.txt
done
