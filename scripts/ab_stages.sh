#!/bin/bash
# A/B of library builds (scratch_libs/*.so vs the in-tree one): stage times of short C2 / C3 runs + C4 scoring rate
mkdir -p gpurun_out; : > gpurun_out/ab_stages.txt
for lib in "" scratch_libs/*.so; do
  for wl in c2:1000 c3:400; do
    SFMGPU_LIB=${lib:+$PWD/$lib} python bench.py --workload ${wl%%:*} --frames ${wl##*:} --steps 3 --warmup 2 --no-cpu-baseline --no-shim --no-c2 --no-c5 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());s=d['stages_ms'];print('${lib:-in-tree} $wl step', round(d['ms_per_step'],2), 'score', round(s['corner_score'],2), 'select', round(s['corner_select'],2), 'ransac', round(s['ransac'],2), 'c4 G/s', round(d['ransac']['value']/1e9,1))" | tee -a gpurun_out/ab_stages.txt
  done
done
