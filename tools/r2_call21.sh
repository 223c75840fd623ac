#!/bin/bash
# Round-2 GPU call 21: k_fpm as the default forward projector: projector / oracle / golden tests, headline and c2 bench
# lines, ncu --set full of k_fpm<4> at the headline size (L2 -> SM bytes per update, after; "before" = ncu_fpq_headline)
set -u
mkdir -p gpurun_out /tmp/rep
timeout 1500 python -m pytest tests/test_gpu_projector.py tests/test_zz_full_size_vs_oracle.py tests/test_gpu_goldens.py tests/test_gpu_goldens_ir.py tests/test_gpu_loops_vs_oracle.py tests/test_gpu_robust_terms.py tests/test_gpu_host_entry_points.py -x -q -m gpu > gpurun_out/r2c21_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2c21_tests.log
timeout 900 python bench.py > gpurun_out/r2c21_bench_headline.json 2> gpurun_out/r2c21_bench_headline.err
echo "bench rc=$?"; cat gpurun_out/r2c21_bench_headline.json
timeout 600 python bench.py --config c2 > gpurun_out/r2c21_bench_c2.json 2> gpurun_out/r2c21_bench_c2.err
echo "bench c2 rc=$?"; cat gpurun_out/r2c21_bench_c2.json
cat > /tmp/prof_fp.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
from tomobar_b200._lib import lib
from tomobar_b200.projector import ProjTools3D
nz, n, na, os_n = 512, 2048, 1800, 24
P = ProjTools3D(n, 0, nz, np.linspace(0, np.pi, na, endpoint=False).astype(np.float32), 0.0, n, "gpu", 0, os_n)
vol = torch.rand((nz, n, n), device="cuda")
out = P._forwprojOSCuPy(vol, 1); torch.cuda.synchronize()
PY
timeout 300 ncu --set full --clock-control none -k regex:k_fpm -c 1 -o /tmp/rep/fpm4 -f python /tmp/prof_fp.py > gpurun_out/r2c21_ncu_fpm.log 2>&1
ncu -i /tmp/rep/fpm4.ncu-rep --page raw --csv > gpurun_out/ncu_fpm4_headline_r02_raw.csv 2>/dev/null
ls -la gpurun_out/ncu_fpm4_headline_r02_raw.csv
