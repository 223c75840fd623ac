#!/bin/bash
# Round-2 GPU call 13 (8 GPUs): headline bench at N = 8 (pairs over peer memory, all-gather inside e2e)
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 24 --warmup 3 > gpurun_out/r2c13_bench_n8.json 2> gpurun_out/r2c13_bench_n8.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2c13_bench_n8.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c13_bench_n8.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"], d["roofline"]["ms_per_launch"], d["kernels"])
PY
