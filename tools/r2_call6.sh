#!/bin/bash
# Round-2 GPU call 6: diagnostic variants of the fused PD_TV kernel (memory-only / L2-resident) beside the real ones
set -u
mkdir -p gpurun_out
timeout 300 python -u tools/check_f2.py 256 1024 512 2048 > gpurun_out/r2c6_check_f2.log 2>&1
grep "PD_TV" gpurun_out/r2c6_check_f2.log
