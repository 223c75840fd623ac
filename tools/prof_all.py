"""One launch of every hot kernel at BASELINE.json's config-2 size (1024 x 1024 x 256, 900 angles,
OS = 6) for ncu captures:  ncu --set full -k regex:'k_' python tools/prof_all.py"""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tomobar_b200._lib import lib, check  # noqa: E402
from tomobar_b200._tensors import ptr  # noqa: E402
from tomobar_b200.methodsIR_CuPy import RecToolsIRCuPy  # noqa: E402
from tomobar_b200.regularisersCuPy import PD_TV_cupy, ROF_TV_cupy  # noqa: E402

n, nz, na, os_n = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (1024, 256, 900, 6)
angles = np.linspace(0.0, math.radians(179.9), na).astype(np.float32)
rec = RecToolsIRCuPy(n, 0, nz, 0.0, angles, n, 0, os_n)
A = rec.Atools
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand((nz, n, n), device="cuda", generator=g) * 0.02
b = torch.rand((nz, na, n), device="cuda", generator=g)
st = torch.cuda.current_stream().cuda_stream
for rep in range(2):
    sino = A._forwprojOSCuPy(x, 0)            # k_vol_to_int, k_fp
    vol = A._backprojOSCuPy(sino, 0)          # k_sino_to_int, k_bp
    grad = A.grad_data_term(x, b, 1, "LS")    # k_vol_to_int, k_fp (fused residual), k_bp
    out = torch.empty_like(x)
    check(lib.tmb_fista_grad_step(ptr(x), ptr(grad), ptr(out), x.numel(), 1e-4, 1, st), "step")
    PD_TV_cupy(x, 3e-4, 2, 0, 1, 12.0, 0, False, out=out)
    PD_TV_cupy(x, 3e-4, 2, 0, 1, 12.0, 0, True, out=out)
    ROF_TV_cupy(x, 3e-4, 2, 1e-3, 0, False, out=out)
    check(lib.tmb_fista_momentum(ptr(out), ptr(x), ptr(grad), x.numel(), 0.5, st), "momentum")
torch.cuda.synchronize()
print("done")
