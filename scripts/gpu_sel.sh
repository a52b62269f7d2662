#!/bin/bash
# corner selection tests (all select modes) + short C2 / C3 benches with stage times; usage: gpu_sel.sh
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout -k 10 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 600 -x -k "corners or pair_frontend or candidates or tracker" > gpurun_out/sel_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/sel_tests.log)" | tee -a gpurun_out/summary.txt
for wl in c2:1000 c3:400; do
  python bench.py --workload ${wl%%:*} --steps 3 --warmup 2 --frames ${wl##*:} --no-cpu-baseline --no-shim > gpurun_out/sel_bench_${wl%%:*}.json 2> gpurun_out/sel_bench_${wl%%:*}.err
  echo "bench ${wl} rc=$?" | tee -a gpurun_out/summary.txt
  python -c "
import json;d=json.load(open('gpurun_out/sel_bench_${wl%%:*}.json'));print(d['ms_per_step'],d['e2e']['ms_per_step'],d['stages_ms'], d.get('parity_in_bench'))" | tee -a gpurun_out/summary.txt
done
