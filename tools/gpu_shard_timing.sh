#!/bin/bash
# gpurun --gpus N: per-kernel breakdown of the sharded step (PZ_SHARD_TIMING=1, events between the kernels) next to the graph run
set -u
mkdir -p gpurun_out
N=${1:-2}
M=${2:-rw1d}
P=10000000; [ $M = rw1d ] && P=100000000
for tm in 0 1; do
  if [ $tm = 1 ]; then export PZ_SHARD_TIMING=1; else unset PZ_SHARD_TIMING; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --model $M --particles $P --steps 30 --warmup 3 --no-cpu --no-parity --dtype f32 > gpurun_out/shard_timing_${M}_g${N}_$tm.json 2> gpurun_out/shard_timing_${M}_g${N}_$tm.err; echo "rc=$?"
  python - <<PY
import json
for l in open('gpurun_out/shard_timing_${M}_g${N}_$tm.json'):
    if l.startswith('{'):
        d=json.loads(l); print('timing=$tm', d['n_gpus'], 'gpus', round(d['ms_per_step'],4), 'ms/step  e2e', round(d['e2e']['ms_per_step'],4))
PY
  grep "pz shard timing" gpurun_out/shard_timing_${M}_g${N}_$tm.err
done
