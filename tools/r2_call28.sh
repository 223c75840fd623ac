#!/bin/bash
# Round-2 GPU call 28: k_fi_gather_w (default): Fourier tests, ncu --set full at config 4, config-4 bench line
set -u
mkdir -p gpurun_out /tmp/rep
timeout 900 python -m pytest tests/test_gpu_fourier.py tests/test_gpu_goldens.py tests/test_gpu_host_arrays.py tests/test_memory_estimator.py tests/test_zz_full_size_gpu.py -x -q > gpurun_out/r2c28_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r2c28_tests.log
timeout 600 python tools/check_gather.py > gpurun_out/r2c28_check_gather.log 2>&1; grep "n=" gpurun_out/r2c28_check_gather.log | awk '{print $1,$2,$3,$5,$6,$7,$8,$13,$14}'
timeout 400 ncu --set full --clock-control none -k regex:k_fi_gather_w -c 1 -o /tmp/rep/gather_w_c4 -f python tools/prof_fourier.py > gpurun_out/r2c28_ncu_c4.log 2>&1
ncu -i /tmp/rep/gather_w_c4.ncu-rep --page raw --csv > gpurun_out/ncu_gather_w_c4_r02_raw.csv 2>/dev/null
python tools/ncu_traffic.py /tmp/rep/gather_w_c4.ncu-rep 64 4096 > gpurun_out/r2c28_traffic.log 2>&1; tail -12 gpurun_out/r2c28_traffic.log
cp profiles/ncu_traffic_r02.json gpurun_out/ncu_traffic_r02.json
timeout 600 python bench.py --config c4 > gpurun_out/bench_r02b_n1_c4.json 2> gpurun_out/bench_r02b_n1_c4.err; cut -c1-200 gpurun_out/bench_r02b_n1_c4.json
