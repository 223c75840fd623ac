"""Diagnostic: FOURIER_INV with the slice-pair gather on / off, repeated, at a shape given as nz na detX."""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tomobar_b200.methodsDIR_CuPy import RecToolsDIRCuPy  # noqa: E402

nz, na, detX = (int(v) for v in sys.argv[1:4])
g = torch.Generator(device="cuda").manual_seed(nz + na)
d = torch.rand((nz, na, detX), device="cuda", generator=g)
angles = np.linspace(0, math.pi, na, endpoint=False).astype(np.float32)
T = RecToolsDIRCuPy(detX, 0, nz, 0.0, angles, detX, device_projector=0)
res = {}
for pairs in (True, False, True, False):
    T._GATHER_SLICE_PAIRS = pairs
    res.setdefault(pairs, []).append(T.FOURIER_INV(d))
print("pairs run 1 == pairs run 2:", torch.equal(res[True][0], res[True][1]))
print("planar run 1 == planar run 2:", torch.equal(res[False][0], res[False][1]))
a, b = res[True][0], res[False][0]
diff = (a - b).abs()
print("pairs == planar:", torch.equal(a, b), "max diff", float(diff.max()), "of", float(b.abs().max()))
per_slice = diff.amax(dim=(1, 2))
print("slices that differ:", [int(i) for i in torch.nonzero(per_slice > 0).flatten()])
