#!/bin/bash
# Runs on an 8-GPU box: the bench at N=8 and N=4 (torchrun, as the driver launches it), the scheduler C++ test on 8 ranks.
mkdir -p gpurun_out; : > gpurun_out/summary.txt
nvidia-smi -L | wc -l >> gpurun_out/summary.txt; nproc >> gpurun_out/summary.txt
for N in 8 4; do
  timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?" | tee -a gpurun_out/summary.txt
done
timeout -k 10 300 ./tests/sched_test 8 41 > gpurun_out/sched_test8.log 2>&1; echo "sched_test rc=$? $(tail -1 gpurun_out/sched_test8.log)" | tee -a gpurun_out/summary.txt
