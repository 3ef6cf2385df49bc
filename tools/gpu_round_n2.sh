#!/bin/bash
# Two-GPU call: the NCCL sharding test, the N = 2 bench at the driver's size and once at BASELINE configs[4]'s per-GPU size (125 h).
set -u
TAG=${1:-n2}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_shard_gpu.py -m gpu -q 2>&1 | tail -15 ) > gpurun_out/${TAG}_pytest_shard.log
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>gpurun_out/${TAG}_bench_10h.err ) > gpurun_out/${TAG}_bench_10h.json
( timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --hours 125 --steps 1 --warmup 1 --no-cpu-baseline --no-extras 2>gpurun_out/${TAG}_bench_125h.err ) > gpurun_out/${TAG}_bench_125h.json
tail -8 gpurun_out/${TAG}_pytest_shard.log; head -c 1500 gpurun_out/${TAG}_bench_10h.json; echo; head -c 1500 gpurun_out/${TAG}_bench_125h.json; tail -c 500 gpurun_out/${TAG}_bench_125h.err
