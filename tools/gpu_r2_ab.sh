#!/bin/bash
# gpurun (round 2): parity of the trajectory tests, A/B of the library variants under build/variants for the
# dtypes in DTYPES and the models in MODELS, then (NCU=1) one ncu --set full capture per dtype of the rw1d step kernel.
set -u
mkdir -p gpurun_out
show() { python -c "
import json,sys
for l in open('$1'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$2', d['dtype'], round(d['value']/1e9,3),'G/s', round(d['ms_per_step'],4),'ms frac',round(d['roofline']['frac'],4), 'tile ms', round(d['roofline']['kernel_ms'],4))"; }
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "trajectory or outlier or tile_size or golden" > gpurun_out/pytest_ab.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ab.log
for dt in ${DTYPES:-f32 f64}; do
for m in ${MODELS:-rw1d}; do
for lib in default $(ls build/variants/ 2>/dev/null); do
    P=10000000; [ $m = rw1d ] && P=100000000
    if [ $lib = default ]; then unset PZ_LIB_PATH; else export PZ_LIB_PATH=$PWD/build/variants/$lib; fi
    timeout 300 python bench.py --model $m --dtype $dt --particles $P --steps 30 --warmup 3 --no-cpu > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
    show gpurun_out/ab.json "$lib $m"
done
done
done
unset PZ_LIB_PATH
if [ "${NCU:-0}" = 1 ]; then
for dt in f32 f64; do
  CMD="python bench.py --model rw1d --dtype $dt --particles 100000000 --steps 3 --warmup 3 --no-cpu"
  ncu --set full --clock-control none --import-source on -k regex:pf_step_tile -s 4 -c 1 -f -o gpurun_out/prof_rw1d_$dt $CMD > gpurun_out/ncu_full_rw1d_$dt.log 2>&1
  echo "ncu $dt rc=$?"
done
fi
