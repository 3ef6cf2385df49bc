#!/bin/bash
# A/B of the residual-layer modes of the direct kernel (same box, alternating): 128-row tiles (DT1) x TMA residual ring.
# usage: bash tools/gpu_ab3.sh <tag>
set -u
TAG=${1:-ab3}
mkdir -p gpurun_out
{
  for i in 1 2; do
    for cfg in "0 0" "0 1" "1 0" "1 1" "1 3"; do        # (defaults: DIRECT_DT1=1 TMA_EPI=1)
      set -- $cfg
      echo "== resnet DIRECT_DT1=$1 TMA_EPI=$2 (#$i)"; ISS_B200_DIRECT_DT1=$1 ISS_B200_TMA_EPI=$2 timeout 150 python tests/tools/resnet_check.py 3 2>&1 | grep -E "^mode 3|rror|Trace|timed out|libiss" | head -5
    done
  done
  echo "== vbx 3 min, defaults"; timeout 200 python tests/tools/vbx_profile.py 3 2>&1 | grep -E "^K[45]|rror" | head -4
  echo "== default CNN"; timeout 200 python tools/tc_check.py 3 10 2>&1 | grep -E "^mode|rror|timed out"
} > gpurun_out/${TAG}_ab.log 2>&1
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 ) > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_ab.log; tail -3 gpurun_out/${TAG}_pytest.log
