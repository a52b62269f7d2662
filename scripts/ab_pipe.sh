#!/bin/bash
# resident stage pipeline sweep: ab_pipe.sh "<pairs per chunk>"
for c in $1; do
  python bench.py --steps 3 --warmup 2 --no-cpu-baseline --pipe $c 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('pipe $c', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))" | tee -a gpurun_out/ab.txt
done
