#!/bin/bash
# gpurun: compute-sanitizer memcheck + racecheck over small-N steps of every kernel family (1 GPU), and memcheck over the
# single-process multi-device filter when the box has 2 GPUs (its mailbox protocol uses relaxed system-scope accesses).
# Summaries land in gpurun_out/sanitize_*.log; copy them under profiles/.
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout 1500 $CS --tool $tool --print-limit 20 python tools/sanitize_driver.py all > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize driver" gpurun_out/sanitize_$tool.log | tail -3
done
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  PZ_P2P_TIMEOUT_S=0 timeout 1500 $CS --tool memcheck --print-limit 20 python tools/sanitize_driver.py multi > gpurun_out/sanitize_memcheck_multi.log 2>&1
  echo "memcheck multi rc=$?"; grep -E "ERROR SUMMARY|sanitize driver" gpurun_out/sanitize_memcheck_multi.log | tail -3
fi
