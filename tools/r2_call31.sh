#!/bin/bash
# Round-2 GPU call 31: k_fi_gather_w build variants (corner skip on / off, rows of patches linear / centre first) at 64 and
# 16 complex slices per launch; the Nyquist fix of the filter stage
set -u
mkdir -p gpurun_out
for v in skip_lin noskip_lin skip_cf noskip_cf; do
  for nz in 128 32; do
    echo "== $v nz=$nz"
    TMB_LIB=$PWD/build_variants/libtmb_$v.so timeout 300 python tools/check_gather.py 2048 2000 $nz 2>&1 | grep "n=" | awk '{print $5,$6,$7,$8,$13,$14}' | tr '\n' ';'; echo
  done
done 2>&1 | tee gpurun_out/r2c31_variants.log
timeout 600 python -m pytest tests/test_gpu_fourier.py tests/test_gpu_host_entry_points.py -x -q > gpurun_out/r2c31_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2c31_tests.log
timeout 300 python tools/diag_filter_pairs.py 2048 16 2000 > gpurun_out/r2c31_diag.log 2>&1; head -4 gpurun_out/r2c31_diag.log
