#!/bin/bash
# streaming tests + chunk-size sweep of the end-to-end call on C3 (and C2)
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_two_view.py -q -m gpu --timeout 600 -x -k "streaming or pair_frontend" > gpurun_out/stream_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/stream_tests.log)" | tee -a gpurun_out/summary.txt
for ch in 0 40 50 64 80; do
  python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --no-c2 --no-c5 --no-shim --chunk $ch > gpurun_out/chunk_$ch.json 2> gpurun_out/chunk_$ch.err
  python -c "
import json;d=json.load(open('gpurun_out/chunk_$ch.json'));print('c3 chunk', $ch, d['ms_per_step'],d['e2e']['ms_per_step'],d['e2e']['h2d_floor_ms'])" | tee -a gpurun_out/summary.txt
done
for ch in 0 32 50; do
  python bench.py --workload c2 --steps 3 --warmup 1 --no-cpu-baseline --no-c2 --no-c5 --no-shim --chunk $ch > gpurun_out/chunkc2_$ch.json 2> gpurun_out/chunkc2_$ch.err
  python -c "
import json;d=json.load(open('gpurun_out/chunkc2_$ch.json'));print('c2 chunk', $ch, d['ms_per_step'],d['e2e']['ms_per_step'],d['e2e']['h2d_floor_ms'])" | tee -a gpurun_out/summary.txt
done
