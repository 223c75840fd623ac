#!/bin/bash
# Round-2 GPU call 34: FOURIER_INV with the filter table cached on the host: timeline gaps, tests, launch list, bench line
set -u
mkdir -p gpurun_out
timeout 300 python tools/gaps_fourier.py > gpurun_out/r2c34_gaps.log 2>&1; grep -v Warn gpurun_out/r2c34_gaps.log | head -8
timeout 900 python -m pytest tests/test_gpu_fourier.py tests/test_gpu_host_entry_points.py tests/test_gpu_goldens.py tests/test_gpu_host_arrays.py -x -q > gpurun_out/r2c34_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r2c34_tests.log
timeout 300 python tools/ab_filter_pairs.py > gpurun_out/r2c34_ab.log 2>&1; cat gpurun_out/r2c34_ab.log
timeout 600 python bench.py --config c4 > gpurun_out/bench_r02d_n1_c4.json 2> gpurun_out/bench_r02d_n1_c4.err; cut -c1-300 gpurun_out/bench_r02d_n1_c4.json
