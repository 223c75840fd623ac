#!/bin/bash
# Round-2 GPU call 54 (4 GPUs): the headline at N = 2 and N = 4 on the end-of-round tree
set -u
mkdir -p gpurun_out
for n in 2 4; do
  CUDA_VISIBLE_DEVICES=$(seq -s, 0 $((n-1))) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29540+n)) bench.py --gpus $n --steps 24 --warmup 3 > gpurun_out/bench_r02_final_n$n.json 2> gpurun_out/bench_r02_final_n$n.err
  echo "== N=$n rc=$?"; cut -c1-200 gpurun_out/bench_r02_final_n$n.json
done
