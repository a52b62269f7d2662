#!/bin/bash
# A/B of KLT kernel variants on the short bench: ab_klt.sh "<modes>" [frames]
mkdir -p gpurun_out
for m in $1; do
  python bench.py --steps 3 --warmup 2 --frames ${2:-999} --no-cpu-baseline --klt-mode $m 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('klt-mode $m', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), {k:round(x,2) for k,x in d['stages_ms'].items()})" | tee -a gpurun_out/ab.txt
done
