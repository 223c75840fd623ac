#!/bin/bash
# Round-2 GPU call 19: k_fpm (multi-angle forward projector) bit-equality vs k_fpq and timing
set -u
mkdir -p gpurun_out
timeout 900 python tools/check_fpm.py time > gpurun_out/r2c19_check_fpm.log 2>&1
echo "rc=$?"; tail -40 gpurun_out/r2c19_check_fpm.log
