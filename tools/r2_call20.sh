#!/bin/bash
# Round-2 GPU call 20: timing spread of k_fpq / k_fpm, then the ncu launch list of one subset projection per mode
set -u
mkdir -p gpurun_out
timeout 900 python tools/time_fpm.py > gpurun_out/r2c20_time_fpm.log 2>&1
echo "rc=$?"; cat gpurun_out/r2c20_time_fpm.log
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,clocks_throttle_reasons.active --format=csv
