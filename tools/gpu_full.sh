#!/bin/bash
# gpurun: the round's evidence run -- GPU tests, bench of every model, default bench (with cpu baseline),
# reference arm, ncu launch list + full captures of the step kernel for the three linear-Gaussian shapes.
set -u
mkdir -p gpurun_out
show() { python -c "
import json,sys
for l in open('$1'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$1', d['n_gpus'],'gpu', round(d['value']/1e9,3),'G/s', round(d['ms_per_step'],4),'ms frac',round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']/1e9,3), d['roofline'].get('step_kernels_ms'), d.get('clocks'))"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for m in rw1d lg42 lg63; do
  P=10000000; [ $m = rw1d ] && P=100000000
  timeout 300 python bench.py --model $m --particles $P --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err; echo "rc=$?"; show gpurun_out/bench_$m.json; tail -2 gpurun_out/bench_$m.err
done
timeout 600 python bench.py --model mtt --steps 20 --warmup 3 > gpurun_out/bench_mtt.json 2> gpurun_out/bench_mtt.err; echo "mtt rc=$?"; show gpurun_out/bench_mtt.json; tail -2 gpurun_out/bench_mtt.err
echo "== default bench (with cpu baseline)"; timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"; cat gpurun_out/bench_default.json
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?"; cat gpurun_out/bench_reference.json
CMD="python bench.py --particles 100000000 --steps 3 --warmup 3 --no-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_rw1d.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
bash tools/gpu_ncu2.sh rw1d lg42 lg63
