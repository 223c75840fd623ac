#!/bin/bash
# Round-2 GPU call 35: device timelines (bench.py --timeline) of one sub-step of the iterative configurations at N = 1
set -u
mkdir -p gpurun_out
for c in headline c2 c3 c5; do
  timeout 600 python bench.py --config $c --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --timeline gpurun_out/timeline_r02_$c > gpurun_out/r2c35_$c.json 2> gpurun_out/r2c35_$c.err
  echo "== $c rc=$?"; head -6 gpurun_out/timeline_r02_$c.rank0 | cut -c1-200
done
