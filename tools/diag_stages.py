"""Which stage of FOURIER_INV differs between two runs on the same input (nz na detX): the stages of
RecToolsDIRCuPy.FOURIER_INV (default branch) replayed with every intermediate kept."""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tomobar_b200._lib import check, lib  # noqa: E402
from tomobar_b200._tensors import ptr  # noqa: E402
from tomobar_b200.methodsDIR_CuPy import RecToolsDIRCuPy  # noqa: E402

nz, na, n = (int(v) for v in sys.argv[1:4])
g = torch.Generator(device="cuda").manual_seed(nz + na)
d = torch.rand((nz, na, n), device="cuda", generator=g)
angles = np.linspace(0, math.pi, na, endpoint=False).astype(np.float32)
T = RecToolsDIRCuPy(n, 0, nz, 0.0, angles, n, device_projector=0)
st = torch.cuda.current_stream().cuda_stream
nz2 = nz // 2
theta = torch.as_tensor(-angles, dtype=torch.float32, device="cuda")
sorted_theta, sorted_idx = torch.sort(theta)
sorted_idx = sorted_idx.to(torch.int32)
mu = -np.log(1e-4) / (2 * n * n)
m = int(np.ceil(2 * n * 1 / np.pi * np.sqrt(-mu * np.log(1e-4) + (mu * n) * (mu * n) / 4)))


def run():
    keep = {}
    datac = torch.empty((nz2, na, n), dtype=torch.complex64, device="cuda")
    T._fourier_filter(d, n, n, True, 4, "shepp", 1.0, pack_into=datac)
    keep["0 filter"] = datac.clone()
    datac = torch.fft.fft(datac, dim=-1)
    keep["1 fft"] = datac.clone()
    check(lib.tmb_fi_scale_sign(ptr(datac), float(np.float32(4 / n)), n, na, nz2, st), "s")
    keep["2 scale"] = datac.clone()
    fde = torch.empty((nz2, 2 * n, 2 * n), dtype=torch.complex64, device="cuda")
    check(lib.tmb_fi_gather(ptr(datac), ptr(fde), ptr(theta), ptr(sorted_theta), ptr(sorted_idx), m, float(np.float32(mu)),
                            n, na, nz2, st), "g")
    keep["3 gather"] = fde.clone()
    fde = torch.fft.ifft2(fde, dim=(-2, -1), norm="forward")
    keep["4 ifft2"] = fde.clone()
    rec = torch.empty((nz, n, n), device="cuda")
    check(lib.tmb_fi_unpad(ptr(rec), ptr(fde), float(np.float32(mu)), float(np.float32(1.0 / (4.0 * n * n))), na, n // 2 + n // 2
                           + 0, nz, 0, n, nz2, st), "u")
    keep["5 unpad"] = rec.clone()
    return keep


a = run()
junk = torch.empty(12345, device="cuda")
b = run()
for k in a:
    x, y = torch.view_as_real(a[k]) if a[k].is_complex() else a[k], torch.view_as_real(b[k]) if b[k].is_complex() else b[k]
    print(f"{k}: equal = {torch.equal(x, y)}  max diff {float((x - y).abs().max()):.3e} of {float(x.abs().max()):.3e}")
