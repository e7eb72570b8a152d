#!/bin/bash
set -u
mkdir -p gpurun_out
CMD="python bench.py --model mtt --steps 10 --warmup 3"
$CMD > gpurun_out/bench_mtt.json 2> gpurun_out/bench_mtt.err; echo "rc=$?"; cut -c1-700 gpurun_out/bench_mtt.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/launches_mtt.csv $CMD > gpurun_out/ncu_list_mtt.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_mtt.csv')) if len(r)>10 and r[0].isdigit()]
agg={}
for r in rows:
    k=r[4].split('(')[0][:40]; agg.setdefault(k,[]).append(float(r[-1]))
for k,v in agg.items(): print(f"{k:42s} n={len(v):3d} mean={sum(v)/len(v)/1000:8.1f} us")
PY
