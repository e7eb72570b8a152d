#!/bin/bash
set -u
mkdir -p gpurun_out
show() { python -c "
import json,sys
for l in open('$1'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$1', d['n_gpus'],'gpu', round(d['value']/1e9,2),'G/s', round(d['ms_per_step'],4),'ms frac',round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']/1e9,2), d['roofline'].get('step_kernels_ms'))"; }
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
for m in rw1d lg42 lg63; do
  P=10000000; [ $m = rw1d ] && P=100000000
  timeout 300 python bench.py --model $m --particles $P --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err; echo "rc=$?"; show gpurun_out/bench_$m.json; tail -2 gpurun_out/bench_$m.err
done
echo "== rw1d TO=8192"; PZ_TILE_OUT=8192 timeout 300 python bench.py --particles 100000000 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_rw1d_to8192.json 2> gpurun_out/bench_rw1d_to8192.err; show gpurun_out/bench_rw1d_to8192.json
echo "== lg no-kron"; for m in lg42 lg63; do PZ_NO_KRON=1 timeout 300 python bench.py --model $m --particles 10000000 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_${m}_nokron.json 2>/dev/null; show gpurun_out/bench_${m}_nokron.json; done
echo "== rw1d sizes"; for P in 1000000 10000000; do timeout 300 python bench.py --particles $P --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_rw1d_$P.json 2>/dev/null; show gpurun_out/bench_rw1d_$P.json; done
