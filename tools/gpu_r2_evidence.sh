#!/bin/bash
# gpurun (round 2, 1 GPU): the evidence run of the shipped build -- GPU suite, default bench (legs + cpu baseline),
# reference arm, ncu launch lists, and one ncu --set full capture of the step kernel per model / storage type given
# as arguments ("rw1d:f64 rw1d:f32 lg42:f64 lg63:f64 mtt").  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
if [ "${SKIP_TESTS:-0}" != 1 ]; then
echo "== pytest gpu"; timeout 2400 python -m pytest tests -q -m gpu --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -18 gpurun_out/pytest_gpu.log
fi
echo "== default bench"; timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"; cat gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?"; cat gpurun_out/bench_reference.json
for spec in "${@:-rw1d:f64}"; do
  m=${spec%%:*}; dt=${spec##*:}; [ "$dt" = "$m" ] && dt=f64
  P=10000000; [ $m = rw1d ] && P=100000000; [ $m = mtt ] && P=1000000
  K=pf_step_tile; [ $m = mtt ] && K=mtt_step
  CMD="python bench.py --model $m --dtype $dt --particles $P --steps 3 --warmup 3 --no-cpu --no-legs"
  $CMD > gpurun_out/plain_${m}_$dt.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_${m}_$dt.log; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_${m}_$dt.csv $CMD > gpurun_out/ncu_list_${m}_$dt.log 2>&1
  echo "launch list $m $dt rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -f -o gpurun_out/prof_${m}_$dt $CMD > gpurun_out/ncu_full_${m}_$dt.log 2>&1
  echo "full $m $dt rc=$?"
  # summaries are made here: the reports (~37 MB each) exceed what gpurun brings back
  key=${m}_$dt
  if [ $m != mtt ]; then python tools/ncu_buckets.py gpurun_out/prof_${m}_$dt.ncu-rep $P 60 $key > gpurun_out/ncu_${m}_${dt}_summary.txt 2>&1
  else python tools/ncu_buckets.py gpurun_out/prof_${m}_$dt.ncu-rep $P 40 > gpurun_out/ncu_${m}_${dt}_summary.txt 2>&1; fi
  echo "---- warp-state samples per source line" >> gpurun_out/ncu_${m}_${dt}_summary.txt
  python tools/ncu_stalls.py gpurun_out/prof_${m}_$dt.ncu-rep 30 >> gpurun_out/ncu_${m}_${dt}_summary.txt 2>&1
  [ "${KEEP_REP:-}" = "$spec" ] || rm -f gpurun_out/prof_${m}_$dt.ncu-rep
done
cp profiles/traffic.json gpurun_out/traffic.json
ls -la gpurun_out/
