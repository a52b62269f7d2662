#!/bin/bash
# Runs on an 8-GPU box: the bench at N=8 (torchrun, as the driver launches it), the scheduler C++ test on 8 ranks, the gloo / NCCL scheduler tests.
mkdir -p gpurun_out; : > gpurun_out/summary.txt
nvidia-smi -L | wc -l >> gpurun_out/summary.txt; nproc >> gpurun_out/summary.txt
timeout -k 10 300 ./tests/sched_test 8 41 > gpurun_out/sched_test8.log 2>&1; echo "sched_test rc=$? $(tail -1 gpurun_out/sched_test8.log)" | tee -a gpurun_out/summary.txt
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$?" | tee -a gpurun_out/summary.txt
timeout -k 10 600 python -m pytest tests/test_gpu_sched.py -q -m gpu --timeout 300 -x > gpurun_out/test_sched.log 2>&1; echo "test_gpu_sched rc=$? $(tail -1 gpurun_out/test_sched.log)" | tee -a gpurun_out/summary.txt
