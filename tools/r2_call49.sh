#!/bin/bash
# Round-2 GPU call 49 (8 GPUs): the headline at N = 8 on the end-of-round tree, with the device timeline of one sub-step
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 24 --warmup 3 --timeline gpurun_out/timeline_r02_n8 > gpurun_out/bench_r02_final_n8.json 2> gpurun_out/bench_r02_final_n8.err
echo "bench rc=$?"; cut -c1-220 gpurun_out/bench_r02_final_n8.json; head -14 gpurun_out/timeline_r02_n8.rank3 | cut -c1-180
