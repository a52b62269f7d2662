#!/bin/bash
# Runs on the GPU box: what the driver runs at round end (smoke, pytest -m gpu in ONE process, both bench arms).
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout -k 10 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke.log)" | tee -a gpurun_out/summary.txt
timeout -k 10 1500 python -m pytest tests -x -q -m gpu --timeout 600 > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu_all.log)" | tee -a gpurun_out/summary.txt
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?" | tee -a gpurun_out/summary.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
