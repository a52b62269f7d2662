#!/bin/bash
# streaming pipeline timeline of one C3 end-to-end pass (SFMGPU_PIPE_TRACE) + chunk-size sweep
mkdir -p gpurun_out
SFMGPU_PIPE_TRACE=1 python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-c2 --no-c5 --no-shim > gpurun_out/trace_bench.json 2> gpurun_out/trace.err
echo "trace rc=$?"
for ch in 50 75 100 125; do
  python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --no-c2 --no-c5 --no-shim --chunk $ch > gpurun_out/chunk_$ch.json 2> gpurun_out/chunk_$ch.err
  python -c "
import json;d=json.load(open('gpurun_out/chunk_$ch.json'));print($ch, d['ms_per_step'],d['e2e']['ms_per_step'],d['e2e']['h2d_floor_ms'])"
done
