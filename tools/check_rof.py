"""k_rof_tv3d_w with the fast normalised differences (default) against the round-1 arithmetic (hook 3: correctly
rounded square root + IEEE division): relative difference after 30 iterations on a few shapes, then ms per iteration."""
import sys

import torch

sys.path.insert(0, ".")
from tomobar_b200._lib import lib  # noqa: E402
from tomobar_b200.regularisersCuPy import ROF_TV_cupy  # noqa: E402


def run(mode, v, lam, its, tau, half, out=None):
    old = lib.tmb_tv_set_simple_kernels(mode)
    try:
        return ROF_TV_cupy(v, lam, its, tau, 0, half, out=out)
    finally:
        lib.tmb_tv_set_simple_kernels(old)


torch.manual_seed(0)
for shape in ((9, 21, 244), (66, 37, 364), (40, 130, 8), (64, 256, 256)):
    for lam, tau, scale in ((3e-4, 1e-3, 0.02), (0.05, 0.02, 1.0)):
        v = torch.randn(*shape, device="cuda") * scale
        for half in (False, True):
            a = run(0, v, lam, 30, tau, half)
            b = run(3, v, lam, 30, tau, half)
            d = (a - b).abs().max().item() / b.abs().max().item()
            print(f"shape={shape} lambda={lam} tau={tau} half={int(half)}: rel max diff fast vs exact {d:.3e} finite={bool(torch.isfinite(a).all())}", flush=True)
its = 20
for nz, n in ((512, 2048), (256, 1024)):
    v = torch.randn(nz, n, n, device="cuda") * 0.02
    out = torch.empty_like(v)
    for mode, name in ((0, "fast"), (3, "exact")):
        for half in (False, True):
            run(mode, v, 3e-4, its, 1e-3, half, out)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                run(mode, v, 3e-4, its, 1e-3, half, out)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 3 / its
            print(f"ROF_TV {name:5s} half={int(half)} {nz}x{n}x{n}: {ms:7.3f} ms/iter  {12 * v.numel() / ms / 1e6:8.1f} GB/s (12 B/voxel)", flush=True)
    del v, out
    torch.cuda.empty_cache()
