#!/bin/bash
# Round-2 GPU call 51: bench lines of the remaining configurations on the end-of-round tree
set -u
mkdir -p gpurun_out
for c in c1 c2 c3 c5; do
  timeout 900 python bench.py --config $c > gpurun_out/bench_r02_final_$c.json 2> gpurun_out/bench_r02_final_$c.err
  echo "== $c rc=$?"; cut -c1-230 gpurun_out/bench_r02_final_$c.json
done
