#!/bin/bash
# e2e chunk-size sweep on the C2 bench: ab_chunk.sh "<chunk sizes>"
for c in $1; do
  python bench.py --steps 3 --warmup 2 --no-cpu-baseline --chunk $c 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('chunk $c', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))" | tee -a gpurun_out/ab.txt
done
