#!/bin/bash
# final evidence of the round: bucket_select capture (C2 shape), launch lists, then what the driver runs (gpu_full.sh)
mkdir -p gpurun_out; : > gpurun_out/summary.txt
WL=c2 bash scripts/gpu_prof.sh 1000 bucket_select r2_bucket_select
bash scripts/gpu_launches.sh 1000 r2_launches_c2 c2
bash scripts/gpu_launches.sh 400 r2_launches_c3 c3
timeout -k 10 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke.log)" | tee -a gpurun_out/summary.txt
timeout -k 10 1500 python -m pytest tests -x -q -m gpu --timeout 600 > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu_all.log)" | tee -a gpurun_out/summary.txt
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?" | tee -a gpurun_out/summary.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
