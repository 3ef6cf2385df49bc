#!/bin/bash
# One GPU call: A/B timings of whole-network variants, tests, bench, launch lists, ncu captures of the dominant kernels.
# Usage: bash tools/gpu_round.sh <tag> [full|fast] [kernel-regex]
set -u
TAG=${1:-run}; MODE=${2:-fast}; KRE=${3:-conv_gemm_tc4h_kernel}
mkdir -p gpurun_out
{
  timeout 200 python tools/tc_check.py 0 10 | grep -E "^mode"
  echo "== ISS_B200_F16_DIRECT=0 (TMEM-operand slab kernel)"; ISS_B200_F16_DIRECT=0 timeout 200 python tools/tc_check.py 3 10 2>&1 | grep -E "^mode|rror|timed out"
  echo "== default (direct kernel)"; timeout 200 python tools/tc_check.py 3 10 2>&1 | grep -E "^mode|rror|timed out"
  echo "== resnet"; timeout 300 python tests/tools/resnet_check.py 2>&1 | grep -E "^mode|rror|Trace"
  echo "== resnet, 1024 windows x 3"; timeout 300 python tests/tools/resnet_check.py 3 2>&1 | grep -E "^mode 3|rror|Trace"
  echo "== K1"; timeout 300 python tests/tools/k1_check.py 10 2>&1 | grep -E "^fp(64|32)"
} > gpurun_out/${TAG}_ab.log 2>&1
( timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/${TAG}_pytest.log
if [ "$MODE" = full ]; then
  ( time timeout 900 python bench.py 2>gpurun_out/${TAG}_bench.err ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench_time.txt
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
# K1: the first 10-h-class launch of k1_check (behind the 6 accuracy launches)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sidekit_features_kernel -s 6 -c 1 -f -o gpurun_out/${TAG}_prof_k1 \
    python tests/tools/k1_check.py 1 > gpurun_out/${TAG}_ncu_k1.log 2>&1
cat gpurun_out/${TAG}_ab.log; tail -12 gpurun_out/${TAG}_pytest.log; head -c 2500 gpurun_out/${TAG}_bench.json; tail -c 600 gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench_time.txt 2>/dev/null
