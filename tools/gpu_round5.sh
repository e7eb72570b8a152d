#!/bin/bash
set -u
mkdir -p gpurun_out
show() { python -c "
import json,sys
for l in open('$1'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$1', d['n_gpus'],'gpu', round(d['value']/1e9,2),'G/s', round(d['ms_per_step'],4),'ms frac',round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']/1e9,2), d.get('posterior_check'))"; }
echo "== pytest host cli + sharded"; timeout 900 python -m pytest tests/test_host_cli.py tests/test_sharded.py -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== 2-GPU bench"
for m in rw1d lg42 lg63; do
  P=10000000; [ $m = rw1d ] && P=100000000
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --model $m --particles $P --steps 20 --warmup 3 > gpurun_out/bench2_$m.json 2> gpurun_out/bench2_$m.err; echo "rc=$?"; show gpurun_out/bench2_$m.json; grep -v "^W\|^$\|\*\*\*\|OMP_NUM" gpurun_out/bench2_$m.err | tail -4
done
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_ref.json
echo "== sanitizer (memcheck) on the smoke + small parity cases"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or degenerate or run_equals_step or multinomial_reference or read_particles" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -8 gpurun_out/sanitizer.log
