#!/bin/bash
# Round-2 GPU call 48: the reference's goldens through the end-of-round CUDA path; FOURIER_INV goldens (direct methods)
set -u
mkdir -p gpurun_out
timeout 900 python tools/golden_report.py > gpurun_out/golden_report_r02.txt 2>&1; tail -30 gpurun_out/golden_report_r02.txt | cut -c1-200
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/r2c48_fi_golden.txt
import sys, numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from golden_cases import load_scan
from tomobar_b200.methodsDIR_CuPy import RecToolsDIRCuPy
data, angles = load_scan()
R = RecToolsDIRCuPy(DetectorsDimH=160, DetectorsDimH_pad=0, DetectorsDimV=128, CenterRotOffset=0.0, AnglesVec=angles, ObjSize=160, device_projector=0)
rec = R.FOURIER_INV(torch.from_numpy(data).cuda(), data_axes_labels_order=["angles", "detY", "detX"], recon_mask_radius=2.0).cpu().numpy()
print(f"FOURIER_INV golden (tests/test_RecToolsDIRCuPy.py:247-248): min {rec.min():.7f} (pinned -0.0372409)  max {rec.max():.7f} (pinned 0.1035610)")
PY
