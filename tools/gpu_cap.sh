#!/bin/bash
# one ncu --set full capture of the dominant convolution of the tree as it is (stamp for bench.py's roofline.traffic)
mkdir -p gpurun_out
timeout 110 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tc4h_kernel -s 2 -c 1 -f -o gpurun_out/r02x_prof \
    python bench.py --hours 0.2 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02x_ncu_full.log 2>&1
tail -3 gpurun_out/r02x_ncu_full.log
