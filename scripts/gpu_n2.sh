#!/bin/bash
# Runs on a 2-GPU box: scheduler C++ test over NCCL, the bench at N=2 (torchrun), A/B of the scoring kernel builds.
mkdir -p gpurun_out; : > gpurun_out/summary.txt
nvidia-smi -L >> gpurun_out/summary.txt
timeout -k 10 300 ./tests/sched_test 2 21 > gpurun_out/sched_test.log 2>&1; echo "sched_test rc=$? $(tail -1 gpurun_out/sched_test.log)" | tee -a gpurun_out/summary.txt
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?" | tee -a gpurun_out/summary.txt
for lib in scratch_libs/*.so; do SFMGPU_LIB=$PWD/$lib timeout 300 python scripts/ab_ransac.py >> gpurun_out/ab_ransac.txt 2>&1; done
timeout 300 python scripts/ab_ransac.py >> gpurun_out/ab_ransac.txt 2>&1
