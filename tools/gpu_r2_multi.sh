#!/bin/bash
# gpurun --gpus N: sharded / multi-device / tracker / host tests, stamps + per-kernel timing of the sharded exchange, benches
set -u
mkdir -p gpurun_out
N=${1:-2}
timeout 1800 python -m pytest tests/test_sharded.py tests/test_mtt.py tests/test_host_cli.py -q -m gpu --durations=5 > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/pytest_multi.log
bash tools/gpu_stamps.sh 2 rw1d f32
bash tools/gpu_shard_timing.sh 2 rw1d
show() { python -c "
import json,sys
for l in open('$1'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$2', d['dtype'], d['n_gpus'], 'gpus', round(d['value']/1e9,3),'G/s', round(d['ms_per_step'],4),'ms frac',round(d['roofline']['frac'],4), 'tile ms', round(d['roofline']['kernel_ms'],4), d.get('sharded_parity'))"; }
for dt in f64 f32; do
for m in rw1d lg42 lg63; do
  P=10000000; [ $m = rw1d ] && P=100000000
  timeout 300 python bench.py --model $m --dtype $dt --particles $P --steps 30 --warmup 3 --no-cpu --no-legs > gpurun_out/b1_${m}_$dt.json 2> gpurun_out/b1.err || tail -3 gpurun_out/b1.err
  show gpurun_out/b1_${m}_$dt.json "$m"
done
for n in 2 $N; do
[ $n = 2 ] && [ $N != 2 ] && [ $dt = f32 ] && continue
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 --dtype $dt > gpurun_out/bench_g${n}_$dt.json 2> gpurun_out/bench_g${n}_$dt.err; echo "bench g$n $dt rc=$?"
show gpurun_out/bench_g${n}_$dt.json "rw1d"
grep -E "PzError" gpurun_out/bench_g${n}_$dt.err | head -3
done
done
timeout 300 python bench.py --model mtt --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_mtt.json 2> gpurun_out/bench_mtt.err; python -c "
import json
for l in open('gpurun_out/bench_mtt.json'):
    if l.startswith('{'):
        d=json.loads(l); print('mtt', round(d['value']/1e9,3),'G/s', round(d['ms_per_step'],4),'ms frac',round(d['roofline']['frac'],4))"
