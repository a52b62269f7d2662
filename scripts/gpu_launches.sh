#!/bin/bash
# ncu launch list (gpu__time_duration per launch) of a short bench run.  usage: gpu_launches.sh [frames] [outname]
mkdir -p gpurun_out
FR=${1:-300}; OUT=${2:-launches}; WL=${3:-c2}
SMALL="python bench.py --workload $WL --steps 1 --warmup 1 --frames $FR --no-cpu-baseline --no-c2 --no-c5 --no-shim"
$SMALL > gpurun_out/plain_$OUT.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/$OUT.csv $SMALL > gpurun_out/ncu_$OUT.log 2>&1
echo "ncu launches rc=$?" | tee -a gpurun_out/summary.txt
python scripts/launch_summary.py gpurun_out/$OUT.csv | tee gpurun_out/$OUT.txt
