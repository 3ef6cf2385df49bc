#!/bin/bash
# A/B of the residual-epilogue pipelining (same box, alternating) + K1 timing.   usage: bash tools/gpu_ab2.sh <tag>
set -u
TAG=${1:-ab2}
mkdir -p gpurun_out
{
  for i in 1 2 3; do
    echo "== resnet RES_PIPE=1 (#$i)"; timeout 300 python tests/tools/resnet_check.py 3 2>&1 | grep -E "^mode 3|rror|Trace"
    echo "== resnet RES_PIPE=0 (#$i)"; ISS_B200_RES_PIPE=0 timeout 300 python tests/tools/resnet_check.py 3 2>&1 | grep -E "^mode 3|rror|Trace"
  done
  echo "== vbx 3 min, RES_PIPE=1"; timeout 300 python tests/tools/vbx_profile.py 3 2>&1 | grep -E "^K[45]|rror"
  echo "== vbx 3 min, RES_PIPE=0"; ISS_B200_RES_PIPE=0 timeout 300 python tests/tools/vbx_profile.py 3 2>&1 | grep -E "^K[45]|rror"
  echo "== default CNN"; timeout 200 python tools/tc_check.py 3 10 2>&1 | grep -E "^mode|rror|timed out"
} > gpurun_out/${TAG}_ab.log 2>&1
( timeout 600 python tests/tools/k1_check.py 10 2>&1 | grep -E "^fp(64|32):" ) > gpurun_out/${TAG}_k1.log
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 ) > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_ab.log gpurun_out/${TAG}_k1.log; tail -4 gpurun_out/${TAG}_pytest.log
