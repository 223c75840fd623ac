#!/bin/bash
# Round-2 GPU call 11 (2 GPUs): NCCL tests again (PZERO ghost first pair, direct output), headline at N = 2
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_tv_shards.py -q -m gpu --timeout 500 -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/r2c11_multi.log
tail -4 gpurun_out/r2c11_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 24 --warmup 3 > gpurun_out/r2c11_bench_n2.json 2> gpurun_out/r2c11_bench_n2.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2c11_bench_n2.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c11_bench_n2.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["ms_per_launch"], d["kernels"])
PY
