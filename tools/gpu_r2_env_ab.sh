#!/bin/bash
# gpurun (1 GPU): A/B of run-time variants of the SAME library: VARIANTS is a ';'-separated list of environment
# settings ("PZ_PERSIST=0;PZ_PERSIST=4;" -- the empty entry is the default), over DTYPES x MODELS.
set -u
mkdir -p gpurun_out
show() { python -c "
import json,sys
for l in open('$1'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('[$2]', d['config']['model'], d['dtype'], round(d['value']/1e9,3),'G/s', round(d['ms_per_step'],4),'ms frac',round(d['roofline']['frac'],4), 'tile ms', round(d['roofline']['kernel_ms'],4))"; }
if [ "${SKIP_TESTS:-0}" != 1 ]; then
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_ab.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ab.log
fi
IFS=';' read -ra VS <<< "${VARIANTS:-;}"
for dt in ${DTYPES:-f64 f32}; do
for m in ${MODELS:-rw1d}; do
for v in "${VS[@]}" ""; do
    P=10000000; [ $m = rw1d ] && P=100000000
    env $v timeout 300 python bench.py --model $m --dtype $dt --particles $P --steps 30 --warmup 3 --no-cpu --no-legs > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
    show gpurun_out/ab.json "$v"
done
done
done
