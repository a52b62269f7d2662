#!/bin/bash
# Runs on a 2-GPU box: scheduler C++ test over NCCL, the bench at N=2 (torchrun), issue-rate probes.
mkdir -p gpurun_out; : > gpurun_out/summary.txt
nvidia-smi -L >> gpurun_out/summary.txt
timeout -k 10 300 ./tests/sched_test 2 21 > gpurun_out/sched_test.log 2>&1; echo "sched_test rc=$? $(tail -1 gpurun_out/sched_test.log)" | tee -a gpurun_out/summary.txt
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?" | tee -a gpurun_out/summary.txt
timeout 120 ./scripts/ubench_pipes > gpurun_out/ubench_pipes.txt 2>&1; echo "ubench rc=$?" | tee -a gpurun_out/summary.txt
