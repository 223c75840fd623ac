"""Diagnostic: FOURIER_INV called repeatedly on the same input (nz na detX): are the calls bit-identical, and if not, the
first torch.fft call whose input or output differs between two calls."""
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
T._GATHER_SLICE_PAIRS = len(sys.argv) > 4 and sys.argv[4] == "pairs"

log = []
orig = {k: getattr(torch.fft, k) for k in ("fft", "ifft", "ifft2")}


def wrap(name):
    def f(x, *a, **kw):
        y = orig[name](x, *a, **kw)
        log[-1].append((name, tuple(x.shape), x.clone(), y.clone()))
        return y
    return f


for k in orig:
    setattr(torch.fft, k, wrap(k))

# the gather's inputs, captured through the pointers the method passes to the library
from tomobar_b200._lib import lib  # noqa: E402
import tomobar_b200.methodsDIR_CuPy as M  # noqa: E402


class _Raw:
    def __init__(self, p, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (int(p), False), "version": 2}


gin = []


class LibProxy:
    def __getattr__(self, name):
        f = getattr(lib, name)
        if name not in ("tmb_fi_gather", "tmb_fi_gather_pairs"):
            return f

        def g(dc, fde, th, sth, sidx, m, mu, n, nproj, c, st):
            gin.append((torch.as_tensor(_Raw(dc, (c, nproj, n, 2), "<f4"), device="cuda").clone(),
                        torch.as_tensor(_Raw(th, (nproj,), "<f4"), device="cuda").clone(),
                        torch.as_tensor(_Raw(sth, (nproj,), "<f4"), device="cuda").clone(),
                        torch.as_tensor(_Raw(sidx, (nproj,), "<i4"), device="cuda").clone(), m, mu))
            return f(dc, fde, th, sth, sidx, m, mu, n, nproj, c, st)
        return g


M.lib = LibProxy()
outs = []
for _ in range(3):
    log.append([])
    outs.append(T.FOURIER_INV(d))
print("gather calls captured:", len(gin))
for i in (1, 2):
    if len(gin) < 3:
        break
    a, b = gin[0], gin[i]
    print(f"gather inputs call 0 vs {i}: samples {torch.equal(a[0], b[0])} theta {torch.equal(a[1], b[1])} sorted theta "
          f"{torch.equal(a[2], b[2])} sorted idx {torch.equal(a[3], b[3])} m {a[4] == b[4]} mu {a[5] == b[5]}")
    if not torch.equal(a[0], b[0]):
        dd = (a[0] - b[0]).abs()
        print(f"   samples: {int((dd > 0).sum())} of {dd.numel()} differ, max {float(dd.max()):.3e} of {float(a[0].abs().max()):.3e}")
for i in (1, 2):
    print(f"call 0 == call {i}:", torch.equal(outs[0], outs[i]), "max diff", float((outs[0] - outs[i]).abs().max()))
    for (n0, s0, x0, y0), (n1, s1, x1, y1) in zip(log[0], log[i]):
        ex, ey = torch.equal(torch.view_as_real(x0), torch.view_as_real(x1)), torch.equal(torch.view_as_real(y0), torch.view_as_real(y1))
        print(f"   {n0} {s0}: input equal {ex}, output equal {ey}")
        if not ex:
            dx = (torch.view_as_real(x0) - torch.view_as_real(x1)).abs().amax(dim=-1)
            bad = torch.nonzero(dx > 0)
            print(f"      {bad.shape[0]} of {dx.numel()} input elements differ, max {float(dx.max()):.3e} of "
                  f"{float(x0.abs().max()):.3e}; nan in input: {bool(torch.isnan(torch.view_as_real(x0)).any())}")
            print("      slices:", sorted(set(int(v) for v in bad[:, 0]))[:30])
            c = x0.shape[-1] // 2
            r = ((bad[:, 1] - c).double() ** 2 + (bad[:, 2] - c).double() ** 2).sqrt()
            print(f"      distance from the grid centre of the differing points: min {float(r.min()):.1f} max {float(r.max()):.1f} "
                  f"(grid half width {c}); first: {bad[:5].tolist()}")
