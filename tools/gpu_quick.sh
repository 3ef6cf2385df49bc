#!/bin/bash
# Lean GPU call: A/B lines of the whole networks + the GPU test-suite (no ncu, no bench).   usage: bash tools/gpu_quick.sh <tag>
set -u
TAG=${1:-q}
mkdir -p gpurun_out
{
  echo "== default"; timeout 200 python tools/tc_check.py 3 10 2>&1 | grep -E "^mode|rror|timed out"
  echo "== ISS_B200_MMA_ORDER=1"; ISS_B200_MMA_ORDER=1 timeout 200 python tools/tc_check.py 3 10 2>&1 | grep -E "^mode|rror|timed out"
  echo "== default again"; timeout 200 python tools/tc_check.py 3 10 2>&1 | grep -E "^mode|rror|timed out"
  echo "== resnet"; timeout 300 python tests/tools/resnet_check.py 2>&1 | grep -E "^mode 3|rror|Trace"
} > gpurun_out/${TAG}_ab.log 2>&1
( timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 ) > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_ab.log; tail -8 gpurun_out/${TAG}_pytest.log
