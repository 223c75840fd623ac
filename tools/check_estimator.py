"""FOURIER_INV dry-run estimate (DeviceMemStack protocol) against the allocator's measured peak:
python tools/check_estimator.py [nz nproj n]   (default: BASELINE.json config 4)"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from tomobar_b200.methodsDIR_CuPy import RecToolsDIRCuPy  # noqa: E402
from tomobar_b200.supp.memory_estimator_helpers import DeviceMemStack  # noqa: E402

nz, nproj, n = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (128, 2000, 2048)
angles = np.linspace(0, np.pi, nproj, endpoint=False).astype(np.float32)
R = RecToolsDIRCuPy(n, 0, nz, 0.0, angles, n, device_projector=0)
with DeviceMemStack() as st:
    shape = R.FOURIER_INV((nz, nproj, n), data_dtype=np.float32)
data = torch.rand((nz, nproj, n), device="cuda")
R.FOURIER_INV(data)
torch.cuda.synchronize()
torch.cuda.empty_cache()
torch.cuda.reset_peak_memory_stats()
before = torch.cuda.memory_allocated()
out = R.FOURIER_INV(data)
torch.cuda.synchronize()
measured = torch.cuda.max_memory_allocated() - before + data.numel() * 4
print(f"shape {nz}x{nproj}x{n}: estimate {st.highwater / 1e9:.3f} GB, measured {measured / 1e9:.3f} GB, "
      f"ratio {measured / st.highwater:.3f}, result {tuple(out.shape)} == {shape}")
