#!/bin/bash
# the residual-mode test with its report + same-box timings of the release variants.   usage: bash tools/gpu_one.sh <tag>
set -u
TAG=${1:-one}
mkdir -p gpurun_out
( timeout 300 python -m pytest tests -m gpu -q -x -k test_k5_residual 2>&1 | grep -E "MODES|AssertionError|passed|failed" | cut -c1-4000 ) > gpurun_out/${TAG}_pytest.log
{
  echo "== resnet default (release once the words are in registers)"; timeout 100 python tests/tools/resnet_check.py 3 2>&1 | grep -E "^mode 3|rror|Trace"
  echo "== resnet TMA_DBG=1 (release after the stores)"; ISS_B200_TMA_DBG=1 timeout 100 python tests/tools/resnet_check.py 3 2>&1 | grep -E "^mode 3|rror|Trace"
  echo "== resnet TMA_EPI=0"; ISS_B200_TMA_EPI=0 timeout 100 python tests/tools/resnet_check.py 3 2>&1 | grep -E "^mode 3|rror|Trace"
  echo "== resnet default again"; timeout 100 python tests/tools/resnet_check.py 3 2>&1 | grep -E "^mode 3|rror|Trace"
} > gpurun_out/${TAG}_ab.log 2>&1
cat gpurun_out/${TAG}_pytest.log gpurun_out/${TAG}_ab.log
