#!/bin/bash
# gpurun: A/B of library variants under build/variants (same ABI), optional small tiles
set -u
mkdir -p gpurun_out
show() { python -c "
import json,sys
for l in open('$1'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$2', round(d['value']/1e9,3),'G/s', 'tile ms', round(d['roofline']['kernel_ms'],4), 'frac',round(d['roofline']['frac'],4))"; }
for lib in default $(ls build/variants/ 2>/dev/null); do
  for tile in big small; do
    for m in rw1d lg42 lg63; do
      P=10000000; [ $m = rw1d ] && P=100000000
      if [ $lib = default ]; then unset PZ_LIB_PATH; else export PZ_LIB_PATH=$PWD/build/variants/$lib; fi
      if [ $tile = small ]; then export PZ_TILE_OUT=small; else unset PZ_TILE_OUT; fi
      timeout 300 python bench.py --model $m --particles $P --steps 20 --warmup 3 --no-cpu > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
      show gpurun_out/ab.json "$lib $tile $m"
    done
  done
done
