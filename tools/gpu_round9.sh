#!/bin/bash
# gpurun: GPU parity tests, short benches, one ncu capture of rw1d
set -u
bash tools/gpu_round8.sh
bash tools/gpu_ncu2.sh rw1d
