#!/bin/bash
# Round-2 GPU call 1: whole GPU suite with the gated tests on, PD_TV kernel families timed, ncu --set full of the fused kernels at the headline size
set -u
mkdir -p gpurun_out
TMB_TEST_UNVALIDATED=1 timeout 600 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r2c1_tests.log
tail -5 gpurun_out/r2c1_tests.log
timeout 200 python -u tools/check_f2.py 256 1024 512 2048 > gpurun_out/r2c1_check_f2.log 2>&1
grep "PD_TV\|False" gpurun_out/r2c1_check_f2.log
for hook in 5 6; do
  TMB_TV_HOOK=$hook timeout 200 ncu --set full --import-source on --clock-control none -k regex:k_pd_tv3d_f2 -c 1 \
      -o gpurun_out/r2c1_f2_hook$hook -f python tools/prof_tv.py 2048 512 2 > gpurun_out/r2c1_ncu_hook$hook.log 2>&1
done
ls -la gpurun_out/r2c1_*
