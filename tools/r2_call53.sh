#!/bin/bash
# Round-2 GPU call 53: ROF kernel with packed fp32 row arithmetic (variant build) against the shipped library
set -u
mkdir -p gpurun_out
rm -f /tmp/rof_ab.pt
timeout 300 python tools/ab_rof_packed.py shipped 2>&1 | tail -3
TMB_LIB=$PWD/build_variants/libtmb_rofpacked.so timeout 300 python tools/ab_rof_packed.py packed 2>&1 | tail -12
TMB_LIB=$PWD/build_variants/libtmb_rofpacked.so timeout 900 python -m pytest tests/test_gpu_vs_reference_kernels.py tests/test_gpu_tv_shards.py tests/test_zz_full_size_vs_oracle.py -x -q -k "rof or ROF" > gpurun_out/r2c53_tests.log 2>&1
echo "tests (packed) rc=$?"; tail -3 gpurun_out/r2c53_tests.log
