#!/bin/bash
# Runs on the GPU box: parity tests touching the corner stage, C2 bench with the walk kernel, and one ncu capture of it.
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout -k 10 900 python -m pytest tests -q -m gpu --timeout 600 -k "corners or candidates or pair_frontend or tracker or golden or c3_shape" > gpurun_out/pytest_walk.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_walk.log)" | tee -a gpurun_out/summary.txt
timeout -k 10 600 python bench.py --workload c2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 rc=$?" | tee -a gpurun_out/summary.txt
WL=c2 bash scripts/gpu_prof.sh 300 score_walk r2_score_walk
