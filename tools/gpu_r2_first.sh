#!/bin/bash
# gpurun (round 2): GPU parity tests, then f32 / f64 benches of the three LG models, then an A/B of the
# f64 tile / occupancy variants under build/variants (same ABI).
set -u
mkdir -p gpurun_out
show() { python -c "
import json,sys
for l in open('$1'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$2', d['dtype'], round(d['value']/1e9,3),'G/s', round(d['ms_per_step'],4),'ms frac',round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']/1e9,3), d['roofline'].get('step_kernels_ms'))"; }
echo "== pytest gpu"; timeout 2400 python -m pytest tests -x -q -m gpu --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -45 gpurun_out/pytest_gpu.log
for dt in f32 f64; do
for m in rw1d lg42 lg63; do
  P=10000000; [ $m = rw1d ] && P=100000000
  timeout 300 python bench.py --model $m --dtype $dt --particles $P --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_${m}_$dt.json 2> gpurun_out/bench_${m}_$dt.err; echo "rc=$?"; show gpurun_out/bench_${m}_$dt.json "$m"; tail -3 gpurun_out/bench_${m}_$dt.err
done
done
for lib in $(ls build/variants/ 2>/dev/null); do
  export PZ_LIB_PATH=$PWD/build/variants/$lib
  timeout 300 python bench.py --model rw1d --dtype f64 --particles 100000000 --steps 20 --warmup 3 --no-cpu > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  show gpurun_out/ab.json "$lib rw1d"
done
