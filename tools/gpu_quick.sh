#!/bin/bash
# gpurun: GPU parity tests (stop at first failure, verbose tail), then short benches.
set -u
mkdir -p gpurun_out
show() { python -c "
import json,sys
for l in open('$1'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$1', d['n_gpus'],'gpu', round(d['value']/1e9,3),'G/s', round(d['ms_per_step'],4),'ms frac',round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']/1e9,3), d['roofline'].get('step_kernels_ms'))"; }
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
for m in rw1d lg42 lg63; do
  P=10000000; [ $m = rw1d ] && P=100000000
  timeout 300 python bench.py --model $m --particles $P --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err; echo "rc=$?"; show gpurun_out/bench_$m.json; tail -3 gpurun_out/bench_$m.err
done
