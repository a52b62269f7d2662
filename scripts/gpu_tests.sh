#!/bin/bash
# Runs on the GPU box: smoke + pytest -m gpu (one process), optional "-k expr" as $1.  Logs go to gpurun_out/.
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout -k 10 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke.log)" | tee -a gpurun_out/summary.txt
if [ -n "$1" ]; then
  timeout -k 10 1500 python -m pytest tests -q -m gpu --timeout 600 -k "$1" > gpurun_out/pytest_gpu.log 2>&1
else
  timeout -k 10 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
fi
echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu.log)" | tee -a gpurun_out/summary.txt
