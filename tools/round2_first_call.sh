#!/bin/bash
# First GPU call of round 2 (one box, ~3 min): everything that was written after round 1's GPU budget ended.
#   gpurun --timeout 420 -- 'bash tools/round2_first_call.sh'
# 1. the gated tests (fused kernel hooks 7 .. 10, tmb_pd_tv_iter2 on one GPU, host-array wrappers, memory estimate)
# 2. timing of every PD_TV kernel family at 1024^2 x 256 and 2048^2 x 512 (hooks 3, 5 .. 10)
# 3. ncu --set full of the fused kernels (hooks 5 and 6) at 1024^2 x 256: stall reasons, DRAM traffic
set -u
mkdir -p gpurun_out
TMB_TEST_UNVALIDATED=1 timeout 150 python -m pytest tests/test_gpu_tv.py tests/test_gpu_tv_shards.py \
    tests/test_gpu_host_arrays.py -q -m gpu --timeout 100 2>&1 | tail -15 | tee gpurun_out/r2_gated_tests.log
timeout 150 python -u tools/check_f2.py 256 1024 512 2048 2>&1 | tee gpurun_out/r2_check_f2.log | grep "PD_TV\|False"
for hook in 5 6; do
  TMB_TV_HOOK=$hook timeout 120 ncu --set full --import-source on --clock-control none -k regex:k_pd_tv3d_f2 -c 2 \
      -o gpurun_out/r2_f2_hook$hook -f python tools/prof_tv.py 1024 256 4 > gpurun_out/r2_ncu_hook$hook.log 2>&1
done
ls -la gpurun_out/r2_*
