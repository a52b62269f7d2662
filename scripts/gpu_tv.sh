#!/bin/bash
# two-view tests (solver modes) + shim tests + a short C2 bench; usage: gpu_tv.sh [frames]
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout -k 10 900 python -m pytest tests/test_gpu_two_view.py tests/test_gpu_shim.py -q -m gpu --timeout 300 -x -s > gpurun_out/tv_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/tv_tests.log)" | tee -a gpurun_out/summary.txt
python bench.py --workload c2 --steps 3 --warmup 2 --frames ${1:-300} --no-cpu-baseline > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
echo "bench rc=$?" | tee -a gpurun_out/summary.txt
python -c "
import json;d=json.load(open('gpurun_out/quick_bench.json'));print(d['ms_per_step'],d['e2e']['ms_per_step'],d['stages_ms']);print({k:v for k,v in d['ransac'].items() if k!='find_E_ransac'});print(d['ransac']['find_E_ransac'])" | tee -a gpurun_out/summary.txt
