#!/bin/bash
# Round-2 GPU call 50: the whole GPU suite once more after the last test edits
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x --timeout 600 -p no:cacheprovider > gpurun_out/r2c50_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r2c50_tests.log
