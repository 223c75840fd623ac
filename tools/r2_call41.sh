#!/bin/bash
# Round-2 GPU call 41: slice-pair gather tests; where two calls of the same FOURIER_INV differ (cuFFT reproducibility)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fourier.py tests/test_gpu_goldens.py tests/test_gpu_host_arrays.py tests/test_memory_estimator.py -x -q > gpurun_out/r2c41_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r2c41_tests.log
timeout 300 python tools/diag_repro.py > gpurun_out/r2c41_repro.log 2>&1; tail -12 gpurun_out/r2c41_repro.log
