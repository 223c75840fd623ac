#!/bin/bash
# Round-2 GPU call 32: FOURIER_INV after the chunk loop, the slice-pair filter, the Nyquist fix and the centre-first gather:
# tests, A/B, estimator, ncu of the gather at the launch shape of the step, launch list, bench line
set -u
mkdir -p gpurun_out /tmp/rep
timeout 900 python -m pytest tests/test_gpu_fourier.py tests/test_gpu_host_entry_points.py tests/test_gpu_goldens.py tests/test_gpu_host_arrays.py tests/test_memory_estimator.py -x -q > gpurun_out/r2c32_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2c32_tests.log
timeout 600 python tools/ab_filter_pairs.py > gpurun_out/r2c32_ab.log 2>&1; cat gpurun_out/r2c32_ab.log
timeout 300 python tools/diag_filter_pairs.py 2048 16 2000 > gpurun_out/r2c32_diag.log 2>&1; head -4 gpurun_out/r2c32_diag.log
timeout 300 python tools/check_estimator.py > gpurun_out/r2c32_estimator.log 2>&1; tail -2 gpurun_out/r2c32_estimator.log
timeout 400 ncu --set full --clock-control none -k regex:k_fi_gather_w -c 1 -o /tmp/rep/gather_w_c4_chunk -f python tools/prof_fourier.py > gpurun_out/r2c32_ncu_c4.log 2>&1
ncu -i /tmp/rep/gather_w_c4_chunk.ncu-rep --page raw --csv > gpurun_out/ncu_gather_w_c4_chunk_r02_raw.csv 2>/dev/null
python tools/ncu_traffic.py /tmp/rep/gather_w_c4_chunk.ncu-rep 16 4096 > gpurun_out/r2c32_traffic.log 2>&1; tail -12 gpurun_out/r2c32_traffic.log
cp profiles/ncu_traffic_r02.json gpurun_out/ncu_traffic_r02.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4_r02.csv python tools/prof_fourier.py > /dev/null 2>&1
timeout 600 python bench.py --config c4 > gpurun_out/bench_r02c_n1_c4.json 2> gpurun_out/bench_r02c_n1_c4.err; cut -c1-300 gpurun_out/bench_r02c_n1_c4.json
