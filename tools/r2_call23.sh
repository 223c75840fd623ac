#!/bin/bash
# Round-2 GPU call 23: ROF warp-strip kernel with the fast normalised differences: accuracy vs the exact path, timing, tests
set -u
mkdir -p gpurun_out
timeout 600 python tools/check_rof.py > gpurun_out/r2c23_check_rof.log 2>&1
echo "rc=$?"; cat gpurun_out/r2c23_check_rof.log
timeout 1500 python -m pytest tests/test_gpu_tv.py tests/test_gpu_vs_reference_kernels.py tests/test_gpu_tv_shards.py tests/test_gpu_regularisers_goldens.py tests/test_gpu_goldens_ir.py tests/test_zz_full_size_vs_oracle.py tests/test_zz_full_size_gpu.py -x -q -m gpu > gpurun_out/r2c23_tests.log 2>&1
echo "tests rc=$?"; tail -8 gpurun_out/r2c23_tests.log
