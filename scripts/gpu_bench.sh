#!/bin/bash
# Runs on the GPU box: default bench, reference arm, then the ncu launch list and one full capture of the KLT kernel.
mkdir -p gpurun_out
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?" | tee -a gpurun_out/summary.txt
SMALL="python bench.py --steps 1 --warmup 1 --frames 120 --no-cpu-baseline"
$SMALL > gpurun_out/plain_small.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?" | tee -a gpurun_out/summary.txt
$SMALL > gpurun_out/plain_small2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:klt_kernel -c 1 -f -o gpurun_out/prof_klt $SMALL > gpurun_out/ncu_klt.log 2>&1
echo "ncu klt rc=$?" | tee -a gpurun_out/summary.txt
