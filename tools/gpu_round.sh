#!/bin/bash
# One GPU call: tests, bench, launch list, ncu capture of the dominant kernel.  Usage: bash tools/gpu_round.sh <tag> [kernel-regex]
set -u
TAG=${1:-run}; KRE=${2:-conv_gemm_tc3h_kernel}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/${TAG}_pytest.log
( timeout 600 python bench.py --steps 2 --warmup 3 2>gpurun_out/${TAG}_bench.err ) > gpurun_out/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --hours 0.5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 6 -c 2 -f -o gpurun_out/${TAG}_prof \
    python bench.py --hours 0.5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
