"""Diagnostic: FOURIER_INV's filter stage (STEP 0) at a multi-chunk size, both paths, against a plain-torch float64
per-slice filter; errors per chunk of slice pairs."""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tomobar_b200.fourier import calc_filter  # noqa: E402
from tomobar_b200.methodsDIR_CuPy import RecToolsDIRCuPy  # noqa: E402

n, nz, na = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (2048, 32, 2000)
angles = np.linspace(0.0, math.radians(179.9), na).astype(np.float32)
R = RecToolsDIRCuPy(n, 0, nz, 0.0, angles, n, device_projector=0)
g = torch.Generator(device="cuda").manual_seed(0)
data = torch.rand((nz, na, n), device="cuda", generator=g)
over = 2 ** math.ceil(math.log2(n * 3))
pm = over // 2 - n // 2
w64 = torch.as_tensor(calc_filter(over, "shepp", 1.0), device="cuda").double() * torch.exp(
    (-2 * np.pi * 1j * 0.5) * torch.fft.rfftfreq(over, device="cuda").double())
torch.view_as_real(w64)[over // 2, 1] = 0.0  # numpy's irfft semantics: the imaginary part of the Nyquist bin is ignored
sgn = torch.where(torch.arange(n, device="cuda") % 2 == 1, 1.0, -1.0)
out = {}
for pairs in (True, False):
    R._FILTER_SLICE_PAIRS = pairs
    datac = torch.zeros((nz // 2, na, n), dtype=torch.complex64, device="cuda")
    R._fourier_filter(data, n, n, True, 4, "shepp", 1.0, pack_into=datac)
    out[pairs] = datac
print("pairs vs per-slice: rel-L2", float((out[True] - out[False]).norm() / out[False].norm()))
for t in range(nz // 2):
    x = torch.nn.functional.pad(data[2 * t:2 * t + 2].double(), (pm, over - pm - n), mode="replicate")
    y = torch.fft.irfft(w64 * torch.fft.rfft(x, dim=2), n=over, dim=2)[:, :, over // 2 - n // 2:over // 2 + n // 2]
    ref = torch.complex(y[0] * sgn, y[1] * sgn)
    e = [float((out[p][t] - ref).norm() / ref.norm()) for p in (True, False)]
    print(f"pair {t:3d}: pairs path {e[0]:.3e}   per-slice path {e[1]:.3e}")
