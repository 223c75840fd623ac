#!/bin/bash
# Round-2 GPU call 5: whole GPU suite with every gate removed (new defaults: f2s/p0 pairs; full-size oracle slices),
# then the PD_TV kernel families side by side incl. the bulk-L2-prefetch variant (hooks 10 / 13)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider -x 2>&1 | tail -40 > gpurun_out/r2c5_tests.log
tail -8 gpurun_out/r2c5_tests.log
timeout 300 python -u tools/check_f2.py 256 1024 512 2048 > gpurun_out/r2c5_check_f2.log 2>&1
grep "PD_TV\|False" gpurun_out/r2c5_check_f2.log
