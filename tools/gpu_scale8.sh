#!/bin/bash
# gpurun --gpus 8: N = 10^9 particles over 8 GPUs (BASELINE.json configs[4]: 1.25e8 per GPU), peer-memory exchange
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for m in ${1:-rw1d lg42}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --model $m --particles 125000000 --steps 30 --warmup 3 > gpurun_out/bench_${m}_g8_1e9.json 2> gpurun_out/bench_${m}_g8_1e9.err; echo "rc=$?"
  python - <<PY
import json
for l in open('gpurun_out/bench_${m}_g8_1e9.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$m', d['n_gpus'], 'gpus', round(d['value']/1e9,2), 'G/s', round(d['ms_per_step'],4), 'ms/step  e2e', round(d['e2e']['ms_per_step'],4), d['posterior_check'])
PY
  grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_${m}_g8_1e9.err | tail -3
done
