#!/bin/bash
# K1 call: accuracy + throughput of the feature kernel, then the GPU tests that depend on it.   usage: bash tools/gpu_k1.sh <tag>
set -u
TAG=${1:-k1}
mkdir -p gpurun_out
( timeout 600 python tests/tools/k1_check.py 10 2>&1 | grep -vE "Warning|warn" ) > gpurun_out/${TAG}_k1.log
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_k1.log; tail -6 gpurun_out/${TAG}_pytest.log
