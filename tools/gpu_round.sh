#!/bin/bash
# Runs on the GPU box via gpurun: smoke, GPU parity tests, bench; everything logged to gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
echo "== bench small"; timeout 300 python bench.py --particles 1000000 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_1e6.json 2> gpurun_out/bench_1e6.err; echo "rc=$?"; cat gpurun_out/bench_1e6.json; tail -3 gpurun_out/bench_1e6.err
echo "== bench full"; timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "rc=$?"; cat gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
for m in lg42 lg63; do
  echo "== bench $m"; timeout 300 python bench.py --model $m --particles 10000000 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err; echo "rc=$?"; cat gpurun_out/bench_$m.json; tail -3 gpurun_out/bench_$m.err
done
