#!/bin/bash
# Runs on the GPU box: the new bench, small first, then the default line, then one reference step.
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout -k 10 600 python bench.py --frames 30 --steps 2 --warmup 1 > gpurun_out/bn_small.json 2> gpurun_out/bn_small.err; echo "small rc=$?" | tee -a gpurun_out/summary.txt
timeout -k 10 900 python bench.py > gpurun_out/bn_full.json 2> gpurun_out/bn_full.err; echo "full rc=$?" | tee -a gpurun_out/summary.txt
timeout -k 10 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bn_ref.json 2> gpurun_out/bn_ref.err; echo "ref rc=$?" | tee -a gpurun_out/summary.txt
nproc >> gpurun_out/summary.txt; free -g | head -2 >> gpurun_out/summary.txt
