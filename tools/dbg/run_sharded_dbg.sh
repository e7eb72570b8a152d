#!/bin/bash
set -u
mkdir -p gpurun_out
export PZ_P2P_TIMEOUT_S=5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --particles 1000000 > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err; echo "bench g2 rc=$?"; tail -c 600 gpurun_out/bench_g2.json; grep -E "PzError|wait mask" gpurun_out/bench_g2.err | head -5
timeout 900 python -m pytest tests/test_sharded.py tests/test_host_cli.py tests/test_mtt.py -x -q -m gpu > gpurun_out/pytest_sh.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_sh.log
