#!/bin/bash
# Round-2 GPU call 33: where the device idles inside a FOURIER_INV call (kernel timeline), new filter tests
set -u
mkdir -p gpurun_out
timeout 300 python tools/gaps_fourier.py > gpurun_out/r2c33_gaps.log 2>&1; tail -45 gpurun_out/r2c33_gaps.log
timeout 900 python -m pytest tests/test_gpu_fourier.py tests/test_gpu_host_entry_points.py -x -q > gpurun_out/r2c33_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r2c33_tests.log
