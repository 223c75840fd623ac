"""Fused two-iteration PD_TV kernels (mode 5, its compile-time-split variant, mode 6, and that variant at
four CTAs per SM, mode 7, or with packets two rows ahead, mode 8) against the
strip kernel (mode 3): agreement on a few shapes, then ms / iteration at the given sizes.
usage: python tools/check_f2.py [nz n [nz n ...]]"""
import sys

import torch

sys.path.insert(0, ".")
from tomobar_b200._lib import lib  # noqa: E402
from tomobar_b200.regularisersCuPy import PD_TV_cupy  # noqa: E402


def run(mode, v, its, out=None, nonneg=1, method=0):
    old = lib.tmb_tv_set_simple_kernels(mode)
    try:
        return PD_TV_cupy(v, 3e-4, its, method, nonneg, 12.0, 0, False, out=out)
    finally:
        lib.tmb_tv_set_simple_kernels(old)


def main():
    torch.manual_seed(0)
    for shape in ((9, 21, 244), (66, 37, 364), (130, 64, 128), (5, 9, 124), (40, 130, 8)):
        v = torch.randn(*shape, device="cuda") * 0.05
        for its, nonneg, method in ((2, 1, 0), (7, 0, 0), (4, 1, 1)):
            b = run(3, v, its, None, nonneg, method)
            for mode in (0, 5, 6, 7, 8, 9, 10, 13, 22, 23, 24, 25, 26, 27, 28, 29):
                a = run(mode, v, its, None, nonneg, method)
                d = (a - b).abs().max().item() / b.abs().max().item()
                print(f"mode {mode} shape={shape} its={its} nonneg={nonneg} methodTV={method}: rel max diff {d:.3e} "
                      f"bit-equal={torch.equal(a, b)} finite={bool(torch.isfinite(a).all())}", flush=True)
    args = [int(a) for a in sys.argv[1:]]
    sizes = list(zip(args[0::2], args[1::2])) or [(512, 2048)]
    its = 20
    for nz, n in sizes:
        v = torch.randn(nz, n, n, device="cuda") * 0.02
        out = torch.empty_like(v)
        ref = None
        for mode, name in ((3, "strip-reg"), (5, "fused-2"), (6, "fused-2s"), (7, "fused-2s/4"), (8, "fused-2s/pf2"), (9, "fused-2s/p0"), (10, "fused-2s/l2pf"), (13, "fused-2s/p0+l2pf"), (0, "default"), (22, "tall 8x4w"), (23, "tall 8x4w/pf2"), (24, "tall 6x5w"), (25, "tall 8x2w x4"), (26, "cta 2x2"), (27, "cta 4x1"), (28, "cta 3x1 x4"), (29, "cta 6x1 x2"), (14, "DIAG mem-only"), (15, "DIAG L2-resident")):
            run(mode, v, its, out)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                run(mode, v, its, out)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 3 / its
            if ref is None:
                ref = out.clone()
            print(f"PD_TV {name:9s} {nz}x{n}x{n}: {ms:8.3f} ms/iter  {36 * v.numel() / ms / 1e6:8.1f} GB/s (36 B/voxel/iter)  "
                  f"rel max diff to strips {((out - ref).abs().max() / ref.abs().max()).item():.2e}", flush=True)
        del v, out, ref
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
