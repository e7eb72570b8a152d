#!/bin/bash
# gpurun --gpus N: sharded parity tests + N-GPU bench
set -u
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L
timeout 900 python -m pytest tests/test_sharded.py -x -q -m gpu > gpurun_out/pytest_sharded.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_sharded.log
for m in rw1d lg42; do
  P=10000000; [ $m = rw1d ] && P=100000000
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --model $m --particles $P --steps 20 --warmup 3 > gpurun_out/bench_${m}_g$N.json 2> gpurun_out/bench_${m}_g$N.err; echo "rc=$?"
  tail -2 gpurun_out/bench_${m}_g$N.json | cut -c1-600; tail -3 gpurun_out/bench_${m}_g$N.err
done
