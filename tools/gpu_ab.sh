#!/bin/bash
# gpurun: A/B of library variants under build/variants (same ABI): MODELS env selects the models
set -u
mkdir -p gpurun_out
show() { python -c "
import json,sys
for l in open('$1'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$2', round(d['value']/1e9,3),'G/s', 'tile ms', round(d['roofline']['kernel_ms'],4), 'frac',round(d['roofline']['frac'],4))"; }
for lib in default $(ls build/variants/ 2>/dev/null); do
  for m in ${MODELS:-rw1d lg42 lg63}; do
    P=10000000; [ $m = rw1d ] && P=100000000
    if [ $lib = default ]; then unset PZ_LIB_PATH; else export PZ_LIB_PATH=$PWD/build/variants/$lib; fi
    timeout 300 python bench.py --model $m --particles $P --steps 30 --warmup 3 --no-cpu > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
    show gpurun_out/ab.json "$lib $m"
  done
done
