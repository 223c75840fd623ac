#!/bin/bash
# Round-2 GPU call 40: k_fi_gather_w with packed FFMA2 accumulation: bit-identity, timing, tests, bench line
set -u
mkdir -p gpurun_out /tmp/rep
for nz in 32 128; do timeout 300 python tools/check_gather.py 2048 2000 $nz 2>&1 | grep "n=" | awk '{print $3,$5,$6,$7,$8,$13,$14}' | tr '\n' ';'; echo; done | tee gpurun_out/r2c40_gather.log
timeout 300 python tools/check_gather.py 2>&1 | grep "n=" | awk '{print $1,$3,$5,$6,$7,$8,$13,$14}' > gpurun_out/r2c40_gather_shapes.log; grep -c "0.00e+00" gpurun_out/r2c40_gather_shapes.log; wc -l gpurun_out/r2c40_gather_shapes.log
timeout 900 python -m pytest tests/test_gpu_fourier.py tests/test_gpu_goldens.py tests/test_gpu_host_arrays.py tests/test_memory_estimator.py tests/test_abi.py -x -q > gpurun_out/r2c40_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r2c40_tests.log
timeout 400 ncu --set full --clock-control none -k regex:k_fi_gather_w -c 1 -o /tmp/rep/gather_w_c4_chunk -f python tools/prof_fourier.py > gpurun_out/r2c40_ncu_c4.log 2>&1
ncu -i /tmp/rep/gather_w_c4_chunk.ncu-rep --page raw --csv > gpurun_out/ncu_gather_w_c4_chunk_r02_raw.csv 2>/dev/null
python tools/ncu_traffic.py /tmp/rep/gather_w_c4_chunk.ncu-rep 16 4096 > gpurun_out/r2c40_traffic.log 2>&1; tail -12 gpurun_out/r2c40_traffic.log
cp profiles/ncu_traffic_r02.json gpurun_out/ncu_traffic_r02.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4_r02.csv python tools/prof_fourier.py > /dev/null 2>&1
timeout 600 python bench.py --config c4 > gpurun_out/bench_r02g_n1_c4.json 2> gpurun_out/bench_r02g_n1_c4.err; cut -c1-300 gpurun_out/bench_r02g_n1_c4.json
