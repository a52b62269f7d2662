#!/bin/bash
# ncu --set full captures of named kernels on a short bench run.  usage: gpu_prof.sh <frames> <regex> <outname> [<regex> <outname> ...]
mkdir -p gpurun_out
FR=$1; shift
SMALL="python bench.py --workload ${WL:-c2} --steps 1 --warmup 1 --frames $FR --no-cpu-baseline --no-c2 --no-c5 --no-shim"
while [ $# -ge 2 ]; do
  RX=$1; OUT=$2; shift 2
  $SMALL > gpurun_out/plain_$OUT.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$RX -c ${NCU_COUNT:-1} -f -o gpurun_out/$OUT $SMALL > gpurun_out/ncu_$OUT.log 2>&1
  echo "ncu $OUT rc=$?" | tee -a gpurun_out/summary.txt
done
