#!/bin/bash
# Host topology of the GPU box (NUMA nodes, memory, GPU affinity) and stage-pipeline A/B on C3.
mkdir -p gpurun_out
{ echo "== nodes"; ls /sys/devices/system/node/ 2>&1; for n in /sys/devices/system/node/node*; do echo $n; grep -E "MemTotal|MemFree" $n/meminfo; cat $n/cpulist; done
  echo "== lscpu"; lscpu | grep -iE "numa|socket|model name|^cpu\(s\)"; echo "== affinity"; taskset -p $$; nproc
  echo "== topo"; nvidia-smi topo -m; echo "== numactl"; which numactl && numactl -H; cat /proc/self/status | grep -i allowed; } > gpurun_out/topo.txt 2>&1
for P in 0 70 140; do
  timeout -k 10 600 python bench.py --steps 3 --warmup 2 --pipe $P --no-cpu-baseline --no-c2 --no-c5 > gpurun_out/bench_pipe$P.json 2> gpurun_out/bench_pipe$P.err; echo "pipe $P rc=$?" >> gpurun_out/summary.txt
done
