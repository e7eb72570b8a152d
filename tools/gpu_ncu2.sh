#!/bin/bash
# gpurun: one ncu --set full capture of the step kernel per model given as arguments (default rw1d)
set -u
mkdir -p gpurun_out
for m in "${@:-rw1d}"; do
  P=10000000; [ $m = rw1d ] && P=100000000
  CMD="python bench.py --model $m --particles $P --steps 3 --warmup 3 --no-cpu"
  $CMD > gpurun_out/plain_$m.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$m.log; continue; }
  ncu --set full --clock-control none --import-source on -k regex:pf_step_tile -s 4 -c 1 -f -o gpurun_out/prof_$m $CMD > gpurun_out/ncu_full_$m.log 2>&1
  echo "full $m rc=$?"
done
ls -la gpurun_out/*.ncu-rep
