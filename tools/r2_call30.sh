#!/bin/bash
# Round-2 GPU call 30: filter-stage diagnostic at a multi-chunk size; the gather per launch shape (64 vs 16 complex slices)
set -u
mkdir -p gpurun_out
timeout 300 python tools/diag_filter_pairs.py > gpurun_out/r2c30_diag.log 2>&1; cat gpurun_out/r2c30_diag.log
timeout 600 python tools/check_gather.py 2048 2000 128 > gpurun_out/r2c30_gather128.log 2>&1; grep "n=" gpurun_out/r2c30_gather128.log | awk '{print $1,$2,$3,$5,$6,$7,$8}'
timeout 600 python tools/check_gather.py 2048 2000 32 > gpurun_out/r2c30_gather32.log 2>&1; grep "n=" gpurun_out/r2c30_gather32.log | awk '{print $1,$2,$3,$5,$6,$7,$8}'
