#!/bin/bash
# gpurun --gpus N (round 2): multi-GPU parity tests (sharded LG filters, both hosts; multi-device tracker), then the
# sharded bench lines (weak scaling, 10^8 per GPU for rw1d, 10^7 for the trackers' LG models) with sharded_parity.
set -u
mkdir -p gpurun_out
N=${1:-2}
timeout 1500 python -m pytest tests/test_sharded.py tests/test_mtt.py tests/test_gpu_parity.py -q -m gpu -k "shard or multi or device" --durations=5 > gpurun_out/pytest_multi_g$N.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_multi_g$N.log
show() { python -c "
import json,sys
for l in open('$1'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$2', d['dtype'], d['n_gpus'], 'gpus', round(d['value']/1e9,3),'G/s', round(d['ms_per_step'],4),'ms e2e', round(d['e2e']['value']/1e9,3), d.get('sharded_parity'))"; }
for spec in ${BENCHES:-rw1d:f64 rw1d:f32 lg42:f64}; do
  m=${spec%%:*}; dt=${spec##*:}
  P=10000000; [ $m = rw1d ] && P=100000000
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --model $m --dtype $dt --particles $P --no-cpu > gpurun_out/bench_g${N}_${m}_$dt.json 2> gpurun_out/bench_g${N}_${m}_$dt.err; echo "bench g$N $m $dt rc=$?"
  show gpurun_out/bench_g${N}_${m}_$dt.json "$m"
  grep -E "PzError|Error" gpurun_out/bench_g${N}_${m}_$dt.err | head -3
done
