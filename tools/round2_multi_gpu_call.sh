#!/bin/bash
# Second GPU call of round 2 (N = 2 GPUs, ~4 min):  gpurun --gpus 2 --timeout 420 -- 'bash tools/round2_multi_gpu_call.sh 2'
# 1. the NCCL tests including the fused pairs of PD_TV iterations over peer memory (ShardedPDTV(pairs=True))
# 2. the sharded bench with single iterations and with pairs (one neighbour synchronisation per pair)
set -u
N=${1:-2}
mkdir -p gpurun_out
TMB_TEST_UNVALIDATED=1 timeout 200 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout 180 2>&1 | tail -15 \
    | tee gpurun_out/r2_multi_tests.log
for extra in "" "--tv-pairs"; do
  timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus "$N" --steps 8 --warmup 3 --no-cpu-baseline --no-e2e $extra > "gpurun_out/r2_bench_n${N}${extra}.log" 2>&1
  grep -m1 '^{' "gpurun_out/r2_bench_n${N}${extra}.log" | cut -c1-400
done
