#!/bin/bash
# Round-2 GPU call 27: the whole GPU suite and smoke() on the end-of-round tree
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x --timeout 600 -p no:cacheprovider > gpurun_out/r2c27_tests.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/r2c27_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
