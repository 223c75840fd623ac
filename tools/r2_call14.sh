#!/bin/bash
# Round-2 GPU call 14: ncu launch list of the headline bench command; ncu --set full of the dominant kernel of
# configs 2, 5 (k_pd_tv3d_f2s) and 4 (k_fi_gather), exported to CSV on the box (the reports exceed the 64 MiB limit)
set -u
mkdir -p gpurun_out /tmp/rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 600 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2c14_bench_under_ncu.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches_r02.csv)"
TMB_TV_HOOK=6 timeout 300 ncu --set full --clock-control none -k regex:k_pd_tv3d_f2s -c 1 -o /tmp/rep/f2s_c2 -f python tools/prof_tv.py 1024 256 2 > gpurun_out/r2c14_ncu_c2.log 2>&1
TMB_TV_HOOK=6 timeout 300 ncu --set full --clock-control none -k regex:k_pd_tv3d_f2s -c 1 -o /tmp/rep/f2s_c5 -f python tools/prof_tv.py 1536 384 2 > gpurun_out/r2c14_ncu_c5.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:k_fi_gather -c 1 -o /tmp/rep/gather_c4 -f python tools/prof_fourier.py > gpurun_out/r2c14_ncu_c4.log 2>&1
for r in f2s_c2 f2s_c5 gather_c4; do ncu -i /tmp/rep/$r.ncu-rep --page raw --csv > gpurun_out/ncu_${r}_r02_raw.csv 2>/dev/null; done
ls -la gpurun_out/ /tmp/rep
