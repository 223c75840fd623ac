#!/bin/bash
# Round-2 GPU call 12: whole GPU suite (scatter branches, 2-D class, Student's t, compat goldens), golden report
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/r2c12_tests.log
tail -30 gpurun_out/r2c12_tests.log
timeout 600 python tools/golden_report.py > gpurun_out/r2c12_golden_report.txt 2>&1; tail -5 gpurun_out/r2c12_golden_report.txt
