#!/bin/bash
# gpurun --gpus N: strong-scaling point of BASELINE.json configs[4]: N_total = 10^9 RW1D particles over N GPUs
set -u
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --model rw1d --total-particles 1000000000 --steps 30 --warmup 3 > gpurun_out/bench_rw1d_strong_g$N.json 2> gpurun_out/bench_rw1d_strong_g$N.err; echo "rc=$?"
python - <<PY
import json
for l in open('gpurun_out/bench_rw1d_strong_g$N.json'):
    if l.startswith('{'):
        d=json.loads(l); print('strong', d['n_gpus'], 'gpus', round(d['value']/1e9,2), 'G/s', round(d['ms_per_step'],4), 'ms/step', d['scaling'], d['config']['particles_per_gpu'], d['posterior_check'])
PY
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_rw1d_strong_g$N.err | tail -3
