#!/bin/bash
# gpurun --gpus N: phase stamps of the sharded exchange (debug build with -DPZ_DBG_STAMPS in build/dbg:
# run `make -C probzelus-haskell_b200/csrc stamps` before sending the tree)
set -u
mkdir -p gpurun_out
N=${1:-2}
M=${2:-rw1d}
P=10000000; [ $M = rw1d ] && P=100000000
export PZ_LIB_PATH=$PWD/build/dbg/libpzgpu_stamps.so
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --model $M --particles $P --steps 30 --warmup 3 --no-cpu --no-parity --dtype ${3:-f32} > gpurun_out/stamps_${M}_g${N}.json 2> gpurun_out/stamps_${M}_g${N}.err; echo "rc=$?"
python - <<PY
import json
for l in open('gpurun_out/stamps_${M}_g${N}.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], 'gpus', round(d['ms_per_step'],4), 'ms/step  e2e', round(d['e2e']['ms_per_step'],4))
PY
grep "pz stamps" gpurun_out/stamps_${M}_g${N}.err
