"""Time tmb_pd_tv_iter2 (the GHOST instantiation of the fused PD_TV kernel, what ShardedPDTV launches) on ONE GPU
beside the whole-volume fused kernel:  python tools/bench_iter2.py nz n"""
import sys

import torch

sys.path.insert(0, ".")
from tomobar_b200._lib import lib, check  # noqa: E402
from tomobar_b200._tensors import ptr, stream_ptr  # noqa: E402
from tomobar_b200.regularisersCuPy import PD_TV_cupy  # noqa: E402

nz, n = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda")
U = [torch.randn(nz + 4, n, n, device=dev) * 0.02 for _ in range(2)]
P = [[torch.zeros(nz + 4, n, n, device=dev) for _ in range(3)] for _ in range(2)]
D = torch.randn(nz + 4, n, n, device=dev) * 0.02


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def iter2(lo, hi, a=0):
    b = 1 - a
    check(lib.tmb_pd_tv_iter2(ptr(D[2:]), ptr(U[a][2:]), ptr(U[b][2:]), ptr(P[a][0][2:]), ptr(P[a][1][2:]), ptr(P[a][2][2:]),
                              ptr(P[b][0][2:]), ptr(P[b][1][2:]), ptr(P[b][2][2:]), nz, n, n, 3e-4, 0, 1, 12.0, lo, hi,
                              *([None] * 10), stream_ptr(D)), "iter2")


for lo, hi in ((0, 0), (1, 1)):
    print(f"tmb_pd_tv_iter2 ghost_lo={lo} ghost_hi={hi} {nz}x{n}x{n}: {timed(lambda: iter2(lo, hi)):8.3f} ms per pair-launch", flush=True)
v = D[2:nz + 2].contiguous()
out = torch.empty_like(v)
for mode, name in ((6, "f2s whole volume"), (0, "default (p0 first pass)")):
    old = lib.tmb_tv_set_simple_kernels(mode)
    ms = timed(lambda: PD_TV_cupy(v, 3e-4, 20, 0, 1, 12.0, 0, False, out=out), 3) / 10
    lib.tmb_tv_set_simple_kernels(old)
    print(f"tmb_pd_tv {name} {nz}x{n}x{n}: {ms:8.3f} ms per pair-launch (20 iterations incl. prologue)", flush=True)
