#!/bin/bash
# Round-2 GPU call 16: fused PD_TV on taller strips (hooks 22-25) against the default
set -u
mkdir -p gpurun_out
timeout 900 python tools/check_f2.py 512 2048 256 1024 > gpurun_out/r2c16_check_f2.log 2>&1
echo "rc=$?"; grep -v "^mode" gpurun_out/r2c16_check_f2.log | tail -40; grep "^mode 2[6-9]" gpurun_out/r2c16_check_f2.log | awk '{print $NF, $(NF-1), $(NF-3)}' | sort | uniq -c | head
