#!/bin/bash
# A/B: outputs of the residual layers by per-warp TMA stores (ISS_B200_TMA_EPI=3) against the default (1).
set -u
TAG=${1:-ab4}
mkdir -p gpurun_out
{
  for i in 1 2; do
    for m in 1 3; do
      echo "== resnet TMA_EPI=$m (#$i)"; ISS_B200_TMA_EPI=$m timeout 150 python tests/tools/resnet_check.py 3 2>&1 | grep -E "^mode 3|rror|Trace|timed out|libiss" | head -5
    done
  done
} > gpurun_out/${TAG}_ab.log 2>&1
( ISS_B200_TMA_EPI=3 timeout 600 python -m pytest tests/test_vbx.py -m gpu -q -x 2>&1 | tail -3 ) > gpurun_out/${TAG}_pytest_tma3.log
cat gpurun_out/${TAG}_ab.log; tail -3 gpurun_out/${TAG}_pytest_tma3.log
