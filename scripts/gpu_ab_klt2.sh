#!/bin/bash
# Runs on the GPU box: KLT parity tests, then A/B of the quad-kernel build variants (klt-mode) on C2 and a 400-frame C3.
mkdir -p gpurun_out; : > gpurun_out/ab.txt
timeout -k 10 900 python -m pytest tests -x -q -m gpu -k "klt or tracker or pair_frontend" --timeout 600 > gpurun_out/pytest_klt.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/pytest_klt.log)" | tee -a gpurun_out/ab.txt
for wl in "c2 1000" "c3 400"; do set -- $wl
for m in ${MODES:-0 17 18 0 17 18}; do
  python bench.py --workload $1 --steps 3 --warmup 2 --frames $2 --no-cpu-baseline --no-c2 --no-c5 --no-shim --klt-mode $m 2>gpurun_out/ab_err.txt | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$1 klt-mode $m', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), {k:round(x,2) for k,x in d['stages_ms'].items()})" | tee -a gpurun_out/ab.txt
done; done
