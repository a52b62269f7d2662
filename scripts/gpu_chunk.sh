#!/bin/bash
# e2e chunk-size sweep on C3 (one GPU) + clean launch lists.
mkdir -p gpurun_out; : > gpurun_out/summary.txt
for CH in 50 100 150; do
  timeout -k 10 600 python bench.py --steps 3 --warmup 2 --chunk $CH --no-cpu-baseline --no-c2 --no-c5 --no-shim > gpurun_out/bench_chunk$CH.json 2> gpurun_out/bench_chunk$CH.err; echo "chunk $CH rc=$?" >> gpurun_out/summary.txt
done
bash scripts/gpu_launches.sh 1000 r2_launches_c2 c2
bash scripts/gpu_launches.sh 400 r2_launches_c3 c3
