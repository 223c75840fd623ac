import sys
import numpy as np, torch
sys.path.insert(0, ".")
from tomobar_b200._lib import lib
from tomobar_b200.projector import ProjTools3D
nz, n, nu, na, os_n = 5, 64, 80, 36, 3
g = torch.Generator(device="cuda").manual_seed(nz)
vol = torch.randn((nz, n, n), device="cuda", generator=g)
b = torch.randn((nz, na, nu), device="cuda", generator=g)
w = torch.rand((nz, na, nu), device="cuda", generator=g)
angles = np.linspace(0, np.pi, na, endpoint=False).astype(np.float32)
res = {}
for mode in (2, 5, 6, 7):
    lib.tmb_fp_set_kernel(mode); lib.tmb_fp_set_segment(24)
    try:
        P = ProjTools3D(nu, 0, nz, angles, 0.5, n, "gpu", 0, os_n)
        sub = os_n - 1
        print("mode", mode, "group", lib.tmb_geom_fp_group(P._g, sub))
        res[mode] = (P._forwprojOSCuPy(vol, sub), P.grad_data_term(vol, b, sub, "PWLS", w), P.grad_data_term(vol, b.abs(), sub, "KL"),
                     P._forwprojOSCuPy(vol, sub))
    finally:
        lib.tmb_fp_set_kernel(0); lib.tmb_fp_set_segment(0)
for mode in (5, 6, 7):
    for i, (a, c) in enumerate(zip(res[2], res[mode])):
        d = (a - c).abs()
        idx = torch.nonzero(d > 0)
        print("mode", mode, "item", i, "shape", tuple(a.shape), "ndiff", idx.shape[0], "max", d.max().item(), "nan", int(torch.isnan(c).sum()))
        for r in idx[:12].tolist():
            print("   ", r, a[tuple(r)].item(), c[tuple(r)].item())
