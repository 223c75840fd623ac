#!/bin/bash
# Round-2 GPU call 10 (2 GPUs): NCCL tests of the sharded TV / FISTA / ADMM (pairs default), then the headline at N = 2
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout 500 -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/r2c10_multi.log
tail -6 gpurun_out/r2c10_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 24 --warmup 3 > gpurun_out/r2c10_bench_n2.json 2> gpurun_out/r2c10_bench_n2.err
echo "bench rc=$?"; tail -c 800 gpurun_out/r2c10_bench_n2.err; tail -c 1500 gpurun_out/r2c10_bench_n2.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 24 --warmup 3 --tv-single --no-e2e > gpurun_out/r2c10_bench_n2_single.json 2> gpurun_out/r2c10_bench_n2_single.err
echo "bench single rc=$?"; tail -c 600 gpurun_out/r2c10_bench_n2_single.json
