"""k_pd_tv3d_f2t (TMA-fed fused PD_TV pass, modes 11 / 12) against the strip kernel (mode 3) and the
register-fed fused kernel (mode 6): agreement on a few shapes for every (warps, stages) instantiation, then
ms / iteration at the given sizes.   usage: python tools/check_f2t.py [nz n [nz n ...]]"""
import sys

import torch

sys.path.insert(0, ".")
from tomobar_b200._lib import lib  # noqa: E402
from tomobar_b200.regularisersCuPy import PD_TV_cupy  # noqa: E402

CONFIGS = ((4, 4), (4, 2), (4, 8), (3, 4), (2, 4), (5, 2))


def run(mode, v, its, out=None, nonneg=1, method=0, cfg=None):
    old = lib.tmb_tv_set_simple_kernels(mode)
    oldc = lib.tmb_tv_set_f2t(*cfg) if cfg else None
    try:
        return PD_TV_cupy(v, 3e-4, its, method, nonneg, 12.0, 0, False, out=out)
    finally:
        lib.tmb_tv_set_simple_kernels(old)
        if oldc is not None:
            lib.tmb_tv_set_f2t(oldc // 10, oldc % 10)


def main():
    torch.manual_seed(0)
    bad = 0
    for shape in ((9, 21, 244), (66, 37, 364), (130, 64, 128), (5, 9, 124), (40, 130, 8), (33, 50, 2044), (3, 2, 4)):
        v = torch.randn(*shape, device="cuda") * 0.05
        for its, nonneg, method in ((2, 1, 0), (7, 0, 0), (4, 1, 1)):
            b = run(3, v, its, None, nonneg, method)
            for mode in (11, 12):
                for cfg in CONFIGS:
                    a = run(mode, v, its, None, nonneg, method, cfg)
                    d = (a - b).abs().max().item() / b.abs().max().item()
                    ok = d < 2e-6 and bool(torch.isfinite(a).all())
                    bad += not ok
                    if not ok or cfg == (4, 4):
                        print(f"mode {mode} cfg={cfg} shape={shape} its={its} nonneg={nonneg} methodTV={method}: "
                              f"rel max diff {d:.3e} {'ok' if ok else 'MISMATCH'}", flush=True)
    print("agreement:", "all ok" if bad == 0 else f"{bad} MISMATCHES", flush=True)
    args = [int(a) for a in sys.argv[1:]]
    sizes = list(zip(args[0::2], args[1::2])) or [(512, 2048)]
    its = 20
    for nz, n in sizes:
        v = torch.randn(nz, n, n, device="cuda") * 0.02
        out = torch.empty_like(v)
        ref = None
        variants = [(3, None, "strip-reg"), (6, None, "fused-2s"), (9, None, "fused-2s/p0")]
        variants += [(11, c, f"f2t {c[0]}x{c[1]}") for c in CONFIGS] + [(12, c, f"f2t/p0 {c[0]}x{c[1]}") for c in CONFIGS[:3]]
        for mode, cfg, name in variants:
            run(mode, v, its, out, cfg=cfg)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                run(mode, v, its, out, cfg=cfg)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 3 / its
            if ref is None:
                ref = out.clone()
            print(f"PD_TV {name:14s} {nz}x{n}x{n}: {ms:8.3f} ms/iter  {36 * v.numel() / ms / 1e6:8.1f} GB/s (36 B/voxel/iter)  "
                  f"rel max diff to strips {((out - ref).abs().max() / ref.abs().max()).item():.2e}", flush=True)
        del v, out, ref
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
