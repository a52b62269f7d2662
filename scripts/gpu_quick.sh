#!/bin/bash
# quick GPU loop: selected parity groups + a short bench.  usage: gpu_quick.sh "<pytest -k expr>" [frames]
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$1" --timeout 300 -x -s > gpurun_out/quick_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/quick_tests.log)" | tee gpurun_out/summary.txt
python bench.py --steps 3 --warmup 2 --frames ${2:-300} --no-cpu-baseline > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
echo "bench rc=$?" | tee -a gpurun_out/summary.txt
python -c "
import json;d=json.load(open('gpurun_out/quick_bench.json'));print(d['ms_per_step'],d['e2e']['ms_per_step'],d['stages_ms'])" | tee -a gpurun_out/summary.txt
