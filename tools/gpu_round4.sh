#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu.log
echo "== bench mtt"; timeout 300 python bench.py --model mtt --steps 30 --warmup 20 > gpurun_out/bench_mtt.json 2> gpurun_out/bench_mtt.err; echo "rc=$?"; cat gpurun_out/bench_mtt.json | cut -c1-1500; tail -3 gpurun_out/bench_mtt.err
