#!/bin/bash
# gpurun --gpus N: sharded parity tests + N-GPU bench (peer-memory exchange, then the NCCL exchange for A/B)
set -u
mkdir -p gpurun_out
N=${1:-2}
MODELS=${2:-"rw1d lg42"}
nvidia-smi -L | head -2
timeout 600 python -m pytest tests/test_sharded.py -x -q -m gpu > gpurun_out/pytest_sharded.log 2>&1; echo "pytest (p2p) rc=$?"; tail -4 gpurun_out/pytest_sharded.log
PZ_SHARD_NCCL=1 timeout 600 python -m pytest tests/test_sharded.py -x -q -m gpu > gpurun_out/pytest_sharded_nccl.log 2>&1; echo "pytest (nccl) rc=$?"; tail -4 gpurun_out/pytest_sharded_nccl.log
for mode in p2p nccl; do
  if [ $mode = nccl ]; then export PZ_SHARD_NCCL=1; else unset PZ_SHARD_NCCL; fi
  for m in $MODELS; do
    P=10000000; [ $m = rw1d ] && P=100000000
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --model $m --particles $P --steps 30 --warmup 3 > gpurun_out/bench_${m}_g${N}_$mode.json 2> gpurun_out/bench_${m}_g${N}_$mode.err; echo "rc=$?"
    python - <<PY
import json
for l in open('gpurun_out/bench_${m}_g${N}_$mode.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$mode $m', d['n_gpus'], 'gpus', round(d['value']/1e9,2), 'G/s', round(d['ms_per_step'],4), 'ms/step  e2e', round(d['e2e']['ms_per_step'],4), d['posterior_check'])
PY
    grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_${m}_g${N}_$mode.err | tail -3
  done
done
