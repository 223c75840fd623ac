#!/bin/bash
# Round-2 multi-GPU call (8 GPUs): NCCL tests, headline bench at N = 8 / 4 / 2 with the round-2 kernels
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout 280 2>&1 | tail -5 | tee gpurun_out/r2c26_multi_tests.log
for N in 8 4 2; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 2951$N \
      bench.py --gpus "$N" --steps 12 --warmup 3 --no-cpu-baseline > "gpurun_out/bench_r02b_n${N}.log" 2>&1
  grep -m1 '^{' "gpurun_out/bench_r02b_n${N}.log" > "gpurun_out/bench_r02b_n${N}.json"
  cut -c1-330 "gpurun_out/bench_r02b_n${N}.json"
done
