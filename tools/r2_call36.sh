#!/bin/bash
# Round-2 GPU call 36 (2 GPUs): device timeline of one z-sharded sub-step of the headline; the NCCL / peer-memory tests
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --timeline gpurun_out/timeline_r02_n2 > gpurun_out/r2c36_n2.json 2> gpurun_out/r2c36_n2.err
echo "bench rc=$?"; cut -c1-200 gpurun_out/r2c36_n2.json; head -30 gpurun_out/timeline_r02_n2.rank0 | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2c36_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2c36_tests.log
