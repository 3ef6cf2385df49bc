#!/bin/bash
# One GPU call: micro-tests, A/B timings of whole-network variants, tests, bench, launch list, ncu capture of the dominant kernel.
# Usage: bash tools/gpu_round.sh <tag> [full|fast] [kernel-regex]
set -u
TAG=${1:-run}; MODE=${2:-fast}; KRE=${3:-conv_gemm_tc4h_kernel}
mkdir -p gpurun_out
{
  timeout 200 python tools/tc_check.py 0 10 | grep -E "^mode"
  echo "== ISS_B200_F16_DIRECT=0 (TMEM-operand slab kernel)"; ISS_B200_F16_DIRECT=0 timeout 200 python tools/tc_check.py 3 10 2>&1 | grep -E "^mode|rror|timed out"
  echo "== default (direct kernel)"; timeout 200 python tools/tc_check.py 3 10 2>&1 | grep -E "^mode|rror|timed out"
  echo "== ISS_B200_FUSE_POOL=0"; ISS_B200_FUSE_POOL=0 timeout 200 python tools/tc_check.py 3 10 2>&1 | grep -E "^mode|rror|timed out"
  echo "== resnet"; timeout 300 python tests/tools/resnet_check.py 2>&1 | grep -E "^mode|rror|Trace"
  echo "== resnet ISS_B200_F16_BN=64"; ISS_B200_F16_BN=64 timeout 300 python tests/tools/resnet_check.py 2>&1 | grep -E "^mode 3|rror|Trace"
  echo "== resnet ISS_B200_RES_BATCH=256"; ISS_B200_RES_BATCH=256 timeout 300 python tests/tools/resnet_check.py 2>&1 | grep -E "^mode 3|rror|Trace"
  echo "== resnet ISS_B200_F16_DIRECT=0"; ISS_B200_F16_DIRECT=0 timeout 300 python tests/tools/resnet_check.py 2>&1 | grep -E "^mode 3|rror|Trace"
} > gpurun_out/${TAG}_ab.log 2>&1
( timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/${TAG}_pytest.log
if [ "$MODE" = full ]; then
  ( timeout 900 python bench.py 2>gpurun_out/${TAG}_bench.err ) > gpurun_out/${TAG}_bench.json
else
  ( timeout 600 python bench.py --no-extras --no-cpu-baseline 2>gpurun_out/${TAG}_bench.err ) > gpurun_out/${TAG}_bench.json
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --hours 0.5 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 6 -c 3 -f -o gpurun_out/${TAG}_prof \
    python bench.py --hours 0.5 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_vbx.csv \
    python tests/tools/vbx_profile.py 3 > gpurun_out/${TAG}_ncu_vbx.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tc4h_kernel -s 30 -c 3 -f -o gpurun_out/${TAG}_prof_resnet \
    python tests/tools/vbx_profile.py 1 > gpurun_out/${TAG}_ncu_resnet.log 2>&1
cat gpurun_out/${TAG}_ab.log; tail -12 gpurun_out/${TAG}_pytest.log; head -c 2500 gpurun_out/${TAG}_bench.json; tail -c 600 gpurun_out/${TAG}_bench.err
