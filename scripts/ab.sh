#!/bin/bash
# A/B of an environment knob on the short bench: ab.sh VAR [frames]
mkdir -p gpurun_out
for v in "" "1"; do
  if [ -n "$v" ]; then export $1=$v; else unset $1; fi
  python bench.py --steps 3 --warmup 2 --frames ${2:-999} --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$1=$v', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), {k:round(x,2) for k,x in d['stages_ms'].items()})" | tee -a gpurun_out/ab.txt
done
