#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
echo "== pytest sharded"; timeout 600 python -m pytest tests/test_sharded.py -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
show() { python -c "
import json,sys
for l in open('$1'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$1', d['n_gpus'],'gpu', round(d['value']/1e9,2),'G/s', round(d['ms_per_step'],4),'ms frac',round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']/1e9,2), d['posterior_check'])"; }
echo "== 2-GPU bench"
for m in rw1d lg42; do
  P=10000000; [ $m = rw1d ] && P=100000000
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --model $m --particles $P --steps 20 --warmup 3 > gpurun_out/bench2_$m.json 2> gpurun_out/bench2_$m.err; echo "rc=$?"; show gpurun_out/bench2_$m.json; tail -5 gpurun_out/bench2_$m.err
done
