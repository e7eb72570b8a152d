#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
for m in rw1d lg42 lg63; do
  P=10000000; [ $m = rw1d ] && P=100000000
  echo "== bench $m"; timeout 300 python bench.py --model $m --particles $P --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err; echo "rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_$m.json')); print(d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['roofline']['frac'], d['roofline']['step_kernels_ms'], 'e2e', d['e2e']['value']/1e9)"; tail -3 gpurun_out/bench_$m.err
done
CMD="python bench.py --particles 100000000 --steps 3 --warmup 2 --no-cpu"
$CMD > gpurun_out/plain_rw1d.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pf_step_tile -s 4 -c 1 -f -o gpurun_out/prof_rw1d $CMD > gpurun_out/ncu_full_rw1d.log 2>&1
echo "full rw1d rc=$?"
CMD2="python bench.py --model lg42 --particles 10000000 --steps 3 --warmup 2 --no-cpu"
$CMD2 > gpurun_out/plain_lg42.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pf_step_tile -s 4 -c 1 -f -o gpurun_out/prof_lg42 $CMD2 > gpurun_out/ncu_full_lg42.log 2>&1
echo "full lg42 rc=$?"
