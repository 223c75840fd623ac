"""GPU idle time inside one FOURIER_INV call at config 4 (or n nz nangles): kernel timeline from torch.profiler (CUPTI),
gaps between consecutive device activities, largest first."""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from timeline import device_timeline  # noqa: E402
from tomobar_b200.methodsDIR_CuPy import RecToolsDIRCuPy  # noqa: E402

n, nz, na = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (2048, 128, 2000)
angles = np.linspace(0.0, math.radians(179.9), na).astype(np.float32)
R = RecToolsDIRCuPy(n, 0, nz, 0.0, angles, n, device_projector=0)
data = torch.rand((nz, na, n), device="cuda")
for _ in range(3):
    R.FOURIER_INV(data, filter_type="shepp", cutoff_freq=1.0)
print(device_timeline(lambda: R.FOURIER_INV(data, filter_type="shepp", cutoff_freq=1.0),
                      f"FOURIER_INV {n}x{n}x{nz}, {na} angles"))
