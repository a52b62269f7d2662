#!/bin/bash
# resident stage pipeline sweep on short C3 / C2 runs: ab_pipe2.sh
mkdir -p gpurun_out; : > gpurun_out/ab_pipe.txt
for wl in c3:400 c2:1000; do
for c in 0 50 100 200; do
  python bench.py --workload ${wl%%:*} --frames ${wl##*:} --steps 3 --warmup 2 --no-cpu-baseline --no-shim --no-c2 --no-c5 --pipe $c 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$wl pipe $c', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))" | tee -a gpurun_out/ab_pipe.txt
done
done
