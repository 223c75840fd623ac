#!/bin/bash
# Round-2 GPU call 18: k_fpm bit-equality after the boundary fix, segment sweep, ncu of k_fpm<3> / k_fpq at the headline
set -u
mkdir -p gpurun_out /tmp/rep
timeout 600 python tools/check_fpm.py > gpurun_out/r2c18_check_fpm.log 2>&1
echo "rc=$?"; tail -20 gpurun_out/r2c18_check_fpm.log
timeout 600 python tools/sweep_fpm.py > gpurun_out/r2c18_sweep_fpm.log 2>&1
echo "rc=$?"; cat gpurun_out/r2c18_sweep_fpm.log
cat > /tmp/prof_fp.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
from tomobar_b200._lib import lib
from tomobar_b200.projector import ProjTools3D
mode = int(sys.argv[1])
nz, n, na, os_n = 512, 2048, 1800, 24
lib.tmb_fp_set_kernel(mode)
P = ProjTools3D(n, 0, nz, np.linspace(0, np.pi, na, endpoint=False).astype(np.float32), 0.0, n, "gpu", 0, os_n)
vol = torch.rand((nz, n, n), device="cuda")
out = P._forwprojOSCuPy(vol, 1); torch.cuda.synchronize()
PY
timeout 300 ncu --set full --clock-control none -k regex:k_fpm -c 1 -o /tmp/rep/fpm3 -f python /tmp/prof_fp.py 6 > gpurun_out/r2c18_ncu_fpm.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:k_fpq -c 1 -o /tmp/rep/fpq -f python /tmp/prof_fp.py 2 > gpurun_out/r2c18_ncu_fpq.log 2>&1
for r in fpm3 fpq; do ncu -i /tmp/rep/$r.ncu-rep --page raw --csv > gpurun_out/ncu_${r}_headline_r02_raw.csv 2>/dev/null; done
ls -la gpurun_out/*.csv | tail -3
