#!/bin/bash
# gpurun: GPU tests, then ncu launch list + one full capture of the dominant kernel per model.
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
CMD="python bench.py --particles 100000000 --steps 3 --warmup 2 --no-cpu"
$CMD > gpurun_out/plain_rw1d.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_rw1d.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain_rw1d.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pf_step_tile -s 4 -c 2 -f -o gpurun_out/prof_rw1d $CMD > gpurun_out/ncu_full_rw1d.log 2>&1
echo "full rw1d rc=$?"
CMD2="python bench.py --model lg42 --particles 10000000 --steps 3 --warmup 2 --no-cpu"
$CMD2 > gpurun_out/plain_lg42.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pf_step_tile -s 4 -c 1 -f -o gpurun_out/prof_lg42 $CMD2 > gpurun_out/ncu_full_lg42.log 2>&1
echo "full lg42 rc=$?"
ls -la gpurun_out/
