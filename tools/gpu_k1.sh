#!/bin/bash
# K1 + ResNet call: accuracy + throughput of the feature kernel, ResNet A/B, then the GPU tests.   usage: bash tools/gpu_k1.sh <tag>
set -u
TAG=${1:-k1}
mkdir -p gpurun_out
( timeout 600 python tests/tools/k1_check.py 10 2>&1 | grep -vE "Warning|warn" ) > gpurun_out/${TAG}_k1.log
{
  echo "== resnet"; timeout 300 python tests/tools/resnet_check.py 2>&1 | grep -E "^mode 3|rror|Trace"
  echo "== default"; timeout 200 python tools/tc_check.py 3 10 2>&1 | grep -E "^mode|rror|timed out"
} > gpurun_out/${TAG}_ab.log 2>&1
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) > gpurun_out/${TAG}_pytest.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sidekit_features_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_k1 \
    python tests/tools/k1_check.py 1 > gpurun_out/${TAG}_ncu_k1.log 2>&1
cat gpurun_out/${TAG}_k1.log gpurun_out/${TAG}_ab.log; tail -6 gpurun_out/${TAG}_pytest.log
