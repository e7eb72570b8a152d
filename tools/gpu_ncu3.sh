#!/bin/bash
# gpurun: ncu --set full of the two small kernels of the step (scan, partition)
set -u
mkdir -p gpurun_out
CMD="python bench.py --model rw1d --particles 100000000 --steps 3 --warmup 3 --no-cpu"
ncu --set full --clock-control none --import-source on -k regex:"pf_scan_blocks|pf_partition" -s 6 -c 2 -f -o gpurun_out/prof_small $CMD > gpurun_out/ncu_small.log 2>&1
echo "rc=$?"
