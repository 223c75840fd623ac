#!/bin/bash
# Round-2 GPU call 2: the TMA-fed fused PD_TV kernel (k_pd_tv3d_f2t): agreement, timing of every (warps, stages), ncu
set -u
mkdir -p gpurun_out
timeout 400 python -u tools/check_f2t.py 256 1024 512 2048 > gpurun_out/r2c2_check_f2t.log 2>&1
grep "PD_TV\|MISMATCH\|agreement\|Error\|error" gpurun_out/r2c2_check_f2t.log | tail -50
TMB_TV_HOOK=11 timeout 200 ncu --set full --import-source on --clock-control none -k regex:k_pd_tv3d_f2t -c 1 \
      -o gpurun_out/r2c2_f2t -f python tools/prof_tv.py 2048 512 2 > gpurun_out/r2c2_ncu.log 2>&1
tail -3 gpurun_out/r2c2_ncu.log
