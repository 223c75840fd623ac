#!/bin/bash
# Round-2 GPU call 25: bench lines of every configuration with k_fpm / the fast ROF arithmetic, ncu launch list of the headline
set -u
mkdir -p gpurun_out
for c in headline c1 c2 c3 c4 c5; do
  timeout 900 python bench.py --config $c > gpurun_out/bench_r02b_n1_$c.json 2> gpurun_out/bench_r02b_n1_$c.err
  echo "$c rc=$? $(cut -c1-160 gpurun_out/bench_r02b_n1_$c.json)"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 600 --csv --log-file gpurun_out/launches_r02b.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2c25_bench_under_ncu.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches_r02b.csv)"
