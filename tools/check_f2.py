"""Fused two-iteration PD_TV kernel (mode 5) against the strip kernel (mode 3): agreement on a few
shapes, then ms / iteration at the given size.   usage: python tools/check_f2.py [nz n]"""
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
    for shape in ((9, 21, 244), (66, 37, 364), (130, 64, 128)):
        v = torch.randn(*shape, device="cuda") * 0.05
        for its in (2, 7):
            a, b = run(5, v, its), run(3, v, its)
            d = (a - b).abs().max().item() / b.abs().max().item()
            print(f"shape={shape} its={its}: rel max diff {d:.3e} bit-equal={torch.equal(a, b)} finite={bool(torch.isfinite(a).all())}",
                  flush=True)
    nz, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) >= 3 else (512, 2048)
    v = torch.randn(nz, n, n, device="cuda") * 0.02
    out = torch.empty_like(v)
    its = 20
    for mode, name in ((3, "strip-reg"), (5, "fused-2")):
        run(mode, v, its, out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            run(mode, v, its, out)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3 / its
        print(f"PD_TV {name:9s} {nz}x{n}x{n}: {ms:8.3f} ms/iter  {36 * v.numel() / ms / 1e6:8.1f} GB/s (36 B/voxel/iter)",
              flush=True)
        ref = out.clone() if mode == 3 else ref
    print("headline-size agreement: rel max diff", ((out - ref).abs().max() / ref.abs().max()).item(),
          "bit-equal", torch.equal(out, ref))


if __name__ == "__main__":
    main()
