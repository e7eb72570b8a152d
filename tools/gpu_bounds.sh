#!/bin/bash
# gpurun (1 GPU): the GPU parity suite against the bounds-checked build (make -C probzelus-haskell_b200/csrc bounds): every staged
# slot, gathered head, ancestor, tile entry, track / appearance count and association pick is asserted on the device.
# (compute-sanitizer is closed on this pool: profiles/r02_sanitizer_closed_on_pool.txt.)
set -u
mkdir -p gpurun_out
export PZ_LIB_PATH=$PWD/build/dbg/libpzgpu_bounds.so
[ -f $PZ_LIB_PATH ] || make -C probzelus-haskell_b200/csrc bounds > gpurun_out/make_bounds.log 2>&1
ls -la $PZ_LIB_PATH || exit 1
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_bounds.log 2>&1; echo "pytest (bounds build) rc=$?"; tail -6 gpurun_out/pytest_bounds.log
