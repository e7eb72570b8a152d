#!/bin/bash
# gpurun --gpus 2 (round 2): the whole GPU suite on a 2-GPU box (sharded / multi-device tests included), then 2-rank benches
# (rank 0 is delayed by the parity check's single-GPU filter: the ranks do not run in lock-step) and an A/B of the library
# variants under build/variants
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --durations=10 > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu2.log
for dt in f64 f32; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --dtype $dt > gpurun_out/bench_g2_$dt.json 2> gpurun_out/bench_g2_$dt.err; echo "bench g2 $dt rc=$?"
python - <<PY
import json
for l in open("gpurun_out/bench_g2_$dt.json"):
    if l.startswith("{"):
        d = json.loads(l); print("$dt", d["n_gpus"], "gpus", round(d["value"]/1e9, 2), "G/s", round(d["ms_per_step"], 4), "ms", d.get("sharded_parity"))
PY
grep -E "PzError" gpurun_out/bench_g2_$dt.err | head -3
done
show() { python -c "
import json,sys
for l in open('$1'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$2', d['dtype'], round(d['value']/1e9,3),'G/s', round(d['ms_per_step'],4),'ms frac',round(d['roofline']['frac'],4), 'tile ms', round(d['roofline']['kernel_ms'],4))"; }
for dt in f32 f64; do
for m in ${MODELS:-rw1d lg42}; do
for lib in default $(ls build/variants/ 2>/dev/null); do
    P=10000000; [ $m = rw1d ] && P=100000000
    if [ $lib = default ]; then unset PZ_LIB_PATH; else export PZ_LIB_PATH=$PWD/build/variants/$lib; fi
    timeout 300 python bench.py --model $m --dtype $dt --particles $P --steps 30 --warmup 3 --no-cpu --no-legs > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
    show gpurun_out/ab.json "$lib $m"
done
done
done
