#!/bin/bash
# A/B of library builds (scratch_libs/*.so vs the in-tree one) on the corner-select stage: short C2 / C3 runs + the corner tests
mkdir -p gpurun_out; : > gpurun_out/ab_select.txt
for lib in "" scratch_libs/*.so; do
  for wl in c2:1000 c3:400; do
    SFMGPU_LIB=${lib:+$PWD/$lib} python bench.py --workload ${wl%%:*} --frames ${wl##*:} --steps 3 --warmup 2 --no-cpu-baseline --no-shim --no-c2 --no-c5 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('${lib:-in-tree} $wl step', round(d['ms_per_step'],2), 'select', round(d['stages_ms']['corner_select'],3))" | tee -a gpurun_out/ab_select.txt
  done
done
