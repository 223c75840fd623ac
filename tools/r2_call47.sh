#!/bin/bash
# Round-2 GPU call 47: the whole GPU suite, smoke() and the bench lines (headline, c4) on the end-of-round tree
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x --timeout 600 -p no:cacheprovider > gpurun_out/r2c47_tests.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/r2c47_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_r02_final_headline.json 2> gpurun_out/bench_r02_final_headline.err; cut -c1-260 gpurun_out/bench_r02_final_headline.json
timeout 600 python bench.py --config c4 > gpurun_out/bench_r02_final_c4.json 2> gpurun_out/bench_r02_final_c4.err; cut -c1-260 gpurun_out/bench_r02_final_c4.json
timeout 300 python tools/check_estimator.py 2>&1 | tail -1
