#!/bin/bash
# A/B of the TMA epilogue mode of the residual layers (same box, alternating).   usage: bash tools/gpu_ab3.sh <tag>
set -u
TAG=${1:-ab3}
mkdir -p gpurun_out
{
  for i in 1 2; do
    for m in 0 1 3; do
      echo "== resnet TMA_EPI=$m (#$i)"; ISS_B200_TMA_EPI=$m timeout 150 python tests/tools/resnet_check.py 3 2>&1 | grep -E "^mode 3|rror|Trace|timed out|libiss" | head -5
    done
  done
  echo "== vbx 3 min, TMA_EPI=3"; ISS_B200_TMA_EPI=3 timeout 200 python tests/tools/vbx_profile.py 3 2>&1 | grep -E "^K[45]|rror" | head -4
} > gpurun_out/${TAG}_ab.log 2>&1
( ISS_B200_TMA_EPI=3 timeout 600 python -m pytest tests/test_vbx.py -m gpu -q -x 2>&1 | tail -5 ) > gpurun_out/${TAG}_pytest_tma3.log
cat gpurun_out/${TAG}_ab.log; tail -3 gpurun_out/${TAG}_pytest_tma3.log
