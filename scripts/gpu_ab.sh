#!/bin/bash
# Runs on the GPU box: A/B of library builds (scratch_libs/*.so vs the in-tree one) with scripts/ab_ransac.py, then GPU tests ($1: -k expr).
mkdir -p gpurun_out; : > gpurun_out/ab_ransac.txt
for lib in scratch_libs/*.so; do SFMGPU_LIB=$PWD/$lib timeout 300 python scripts/ab_ransac.py >> gpurun_out/ab_ransac.txt 2>&1; done
timeout 300 python scripts/ab_ransac.py >> gpurun_out/ab_ransac.txt 2>&1
bash scripts/gpu_tests.sh "$1"
