"""FORWPROJ time of k_fpq for several line-segment lengths: python tools/bench_fp_segments.py n nz na"""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tomobar_b200._lib import lib  # noqa: E402
from tomobar_b200.methodsDIR_CuPy import RecToolsDIRCuPy  # noqa: E402

n, nz, na = (int(v) for v in sys.argv[1:4])
angles = np.linspace(0.0, math.radians(179.9), na).astype(np.float32)
vol = torch.rand((nz, n, n), device="cuda")
for seg in (0, 54, 81, 120, 162, 243, 324, 648, 100000):
    lib.tmb_fp_set_segment(seg)
    R = RecToolsDIRCuPy(n, 0, nz, 0.0, angles, n, device_projector=0)
    lib.tmb_fp_set_segment(0)
    R.FORWPROJ(vol)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        R.FORWPROJ(vol)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(f"n={n} nz={nz} na={na} segment={seg:6d} lines: {ms:8.2f} ms  {float(nz) * n * n * na / ms / 1e6:8.1f} GUPS", flush=True)
    del R
    torch.cuda.empty_cache()
