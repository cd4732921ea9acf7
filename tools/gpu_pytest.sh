#!/bin/bash
# run a subset of the GPU tests under gpurun: tools/gpu_pytest.sh "<-k expr>"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -k "$1" > gpurun_out/pytest_sub.log 2>&1
grep -E "^E  |Error|passed|failed" gpurun_out/pytest_sub.log | cut -c1-260 | head -${2:-25}
