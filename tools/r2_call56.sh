#!/bin/bash
# Round-2 GPU call 56: ncu --set full of the round-2 ROF kernel at 2048^2 x 512 (config 3's dominant kernel)
set -u
mkdir -p gpurun_out /tmp/rep
timeout 600 ncu --set full --clock-control none -k regex:k_rof_tv3d_w -s 1 -c 1 -o /tmp/rep/rof_c3 -f python tools/prof_tv_big.py > gpurun_out/r2c56_ncu.log 2>&1
ncu -i /tmp/rep/rof_c3.ncu-rep --page raw --csv > gpurun_out/ncu_rof_c3_r02_raw.csv 2>/dev/null
python tools/ncu_traffic.py /tmp/rep/rof_c3.ncu-rep 512 2048 > gpurun_out/r2c56_traffic.log 2>&1; tail -12 gpurun_out/r2c56_traffic.log
cp profiles/ncu_traffic_r02.json gpurun_out/ncu_traffic_r02.json
