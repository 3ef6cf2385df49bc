#!/bin/bash
# A/B of ResNet101 modes of the direct kernel (same box, alternating) + the GPU test-suite.   usage: bash tools/gpu_ab3.sh <tag>
# switches (defaults): ISS_B200_DIRECT_PAD (1: 'same' 3x3 layers on the direct kernel), ISS_B200_DIRECT_DT1 (1), ISS_B200_TMA_EPI (1)
set -u
TAG=${1:-ab3}
mkdir -p gpurun_out
{
  for i in 1 2; do
    for pad in 0 1; do
      echo "== resnet DIRECT_PAD=$pad (#$i)"; ISS_B200_DIRECT_PAD=$pad timeout 150 python tests/tools/resnet_check.py 3 2>&1 | grep -E "^mode 3|rror|Trace|timed out|libiss" | head -5
    done
  done
  echo "== vbx 3 min, defaults"; timeout 200 python tests/tools/vbx_profile.py 3 2>&1 | grep -E "^K[45]|rror" | head -4
} > gpurun_out/${TAG}_ab.log 2>&1
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 ) > gpurun_out/${TAG}_pytest.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_vbx.csv \
    python tests/tools/vbx_profile.py 3 > gpurun_out/${TAG}_ncu_vbx.log 2>&1
cat gpurun_out/${TAG}_ab.log; tail -6 gpurun_out/${TAG}_pytest.log
