#!/bin/bash
# one GPU test with its full failure message.   usage: bash tools/gpu_one.sh <tag> <pytest -k expression>
set -u
TAG=${1:-one}; K=${2:-test_k5}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q -x -k "$K" 2>&1 | grep -E "MODES|AssertionError|passed|failed" | cut -c1-3000 ) > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_pytest.log
