"""Times one TV iteration (CUDA events) at a given volume size for each kernel family.
usage: python tools/bench_tv.py [nz n] ...   prints GB/s against the 36 / 24 / 40 B/voxel figures
(per ITERATION: the fused-2 kernel moves 18 B/voxel per iteration, so its figure can exceed the HBM peak)."""
import sys

import torch

sys.path.insert(0, ".")
from tomobar_b200._lib import lib  # noqa: E402
from tomobar_b200.regularisersCuPy import PD_TV_cupy, ROF_TV_cupy  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    sizes = [(256, 1024), (512, 2048)]
    if len(sys.argv) >= 3:
        sizes = [(int(sys.argv[1]), int(sys.argv[2]))]
    its = 20
    for nz, n in sizes:
        v = torch.randn(nz, n, n, device="cuda") * 0.02
        out = torch.empty_like(v)
        nvox = v.numel()
        for mode, name in ((10, "fused-2s/l2pf"), (9, "fused-2s/p0"), (8, "fused-2s/pf2"), (7, "fused-2s/4"), (6, "fused-2s"), (5, "fused-2"), (4, "strip-tma"), (3, "strip-reg"), (2, "cta-march")):
            lib.tmb_tv_set_simple_kernels(mode)
            for half in ((False,) if mode >= 5 else (False, True)):
                ms = timed(lambda: PD_TV_cupy(v, 3e-4, its, 0, 1, 12.0, 0, half, out=out)) / its
                bpv = 24 if half else 36
                print(f"PD_TV {name:9s} half={int(half)} {nz}x{n}x{n}: {ms:8.3f} ms/iter  {bpv * nvox / ms / 1e6:8.1f} GB/s",
                      flush=True)
        lib.tmb_tv_set_simple_kernels(0)
        for half in (False, True):
            ms = timed(lambda: ROF_TV_cupy(v, 3e-4, its, 1e-3, 0, half, out=out)) / its
            print(f"ROF_TV fused     half={int(half)} {nz}x{n}x{n}: {ms:8.3f} ms/iter  {12 * nvox / ms / 1e6:8.1f} GB/s (12 B/voxel), "
                  f"{40 * nvox / ms / 1e6:8.1f} GB/s (reference structure 40 B/voxel)", flush=True)
        del v, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
