#!/bin/bash
# Round-2 GPU call 42: the stage at which two runs of FOURIER_INV on the same input part ways (24 / 40 / 16 complex slices)
set -u
mkdir -p gpurun_out
for s in "48 50 64" "80 50 80" "32 64 96"; do echo "== $s"; timeout 300 python tools/diag_stages.py $s 2>&1 | tail -6; done | tee gpurun_out/r2c42_stages.log
