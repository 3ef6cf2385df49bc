#!/bin/bash
# the GPU test-suite + smoke() of the tree as it is (what the driver runs at round end).   usage: bash tools/gpu_one.sh <tag>
set -u
TAG=${1:-one}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -vE "Warning|warn|^$|host = torch" | tail -12 ) > gpurun_out/${TAG}_pytest.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -3 ) > gpurun_out/${TAG}_smoke.log
cat gpurun_out/${TAG}_pytest.log gpurun_out/${TAG}_smoke.log
