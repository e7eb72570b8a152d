#!/bin/bash
# gpurun (1 GPU): the whole GPU suite, then the step-kernel numbers of every model and storage type and the tracker
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -16 gpurun_out/pytest_gpu.log
show() { python -c "
import json,sys
for l in open('$1'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$2', d['dtype'], d['n_gpus'], 'gpus', round(d['value']/1e9,3),'G/s', round(d['ms_per_step'],4),'ms frac',round(d['roofline']['frac'],4), 'tile ms', round(d['roofline']['kernel_ms'],4))"; }
for dt in f64 f32; do
for m in rw1d lg42 lg63; do
  P=10000000; [ $m = rw1d ] && P=100000000
  timeout 300 python bench.py --model $m --dtype $dt --particles $P --steps 30 --warmup 3 --no-cpu --no-legs > gpurun_out/b1_${m}_$dt.json 2> gpurun_out/b1.err || tail -3 gpurun_out/b1.err
  show gpurun_out/b1_${m}_$dt.json "$m"
done
done
timeout 300 python bench.py --model mtt --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_mtt.json 2> gpurun_out/bench_mtt.err; python -c "
import json
for l in open('gpurun_out/bench_mtt.json'):
    if l.startswith('{'):
        d=json.loads(l); print('mtt', round(d['value']/1e9,3),'G/s', round(d['ms_per_step'],4),'ms frac',round(d['roofline']['frac'],4))"
