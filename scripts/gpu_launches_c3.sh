#!/bin/bash
# ncu launch list of the C3 shape (4K, 8000 corners), 120 frames
mkdir -p gpurun_out
SMALL="python bench.py --workload c3 --steps 1 --warmup 1 --frames 120 --no-cpu-baseline"
$SMALL > gpurun_out/plain_c3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3.csv $SMALL > gpurun_out/ncu_c3.log 2>&1
echo "rc=$?"; python scripts/launch_summary.py gpurun_out/launches_c3.csv | head -20
