#!/bin/bash
# First GPU call of the next round: parity + timing of the kernel variants prepared off-line, each in its own
# process under its own timeout (a trapped kernel must not take the others down), then the contention
# micro-benchmark.  Usage (from the repo root, on the GPU box):  bash tools/round2_bringup.sh 2>&1 | tee gpurun_out/bringup.log
set -u
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" 2>&1 | grep -E "^mode|prof|rror|timed out|Traceback" ; }
timeout 120 python tools/tc_check.py 0 10 2>&1 | grep -E "^mode"            # fp32 reference for the error column
run timeout 100 python tools/tc_check.py 2 10                                # validated default
run ISS_B200_TC3_V2=1 timeout 100 python tools/tc_check.py 2 10              # conflict-free swizzle key + ld.shared
run ISS_B200_TC3_V2=1 ISS_B200_FUSE_POOL=1 timeout 100 python tools/tc_check.py 2 10
run ISS_B200_FUSE_POOL=1 timeout 100 python tools/tc_check.py 2 10
run ISS_B200_FIRST_V2=1 timeout 100 python tools/tc_check.py 2 10            # first layer: row inputs kept in registers
run timeout 100 python tools/tc_check.py 3 10                                # fp16-split engine
run ISS_B200_FUSE_POOL=1 timeout 100 python tools/tc_check.py 3 10
( cd tools && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ../gpurun_out/umma_contention_bench umma_contention_bench.cu ) && timeout 120 gpurun_out/umma_contention_bench
