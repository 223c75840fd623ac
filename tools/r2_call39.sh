#!/bin/bash
# Round-2 GPU call 39: diagnostic of the slice-pair gather at 24 complex slices
set -u
mkdir -p gpurun_out
timeout 300 python tools/check_gather.py 80 50 48 2>&1 | grep "n=" | awk '{print $5,$6,$7,$8,$13,$14}' | tr '\n' ';'; echo
timeout 300 python tools/diag_pairs.py 48 50 80 2>&1 | tail -6
timeout 300 python tools/diag_pairs.py 80 50 80 2>&1 | tail -6
timeout 300 python tools/diag_pairs.py 32 64 96 2>&1 | tail -6
