#!/bin/bash
# Round-2 GPU call 55: ncu launch list of the default bench command on the end-of-round tree
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2c55_b.log 2>&1
echo "rc=$?"; wc -l gpurun_out/launches_r02.csv
