"""Tiny driver for ncu captures of the TV kernels: python tools/prof_tv.py [n] [nz] [iters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tomobar_b200.regularisersCuPy import PD_TV_cupy, ROF_TV_cupy

from tomobar_b200._lib import lib

# TMB_TV_HOOK=<0..8> selects the PD_TV kernel family (tmb_tv_set_simple_kernels), e.g. 6 = k_pd_tv3d_f2s
lib.tmb_tv_set_simple_kernels(int(os.environ.get("TMB_TV_HOOK", "0")))
if os.environ.get("TMB_F2T_CFG"):
    lib.tmb_tv_set_f2t(*[int(v) for v in os.environ["TMB_F2T_CFG"].split()])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nz = int(sys.argv[2]) if len(sys.argv) > 2 else 256
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 6
x = torch.rand(nz, n, n, device="cuda") * 0.02
out = torch.empty_like(x)
for half in (False, True):
    for fn, name in ((PD_TV_cupy, "pd"), (ROF_TV_cupy, "rof")):
        args = (x, 3e-4, iters, 0, 1, 12.0, 0, half) if name == "pd" else (x, 3e-4, iters, 1e-3, 0, half)
        fn(*args, out=out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(*args, out=out)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / iters
        nv = x.numel()
        bpv = {("pd", False): 36, ("pd", True): 24, ("rof", False): 12, ("rof", True): 12}[(name, half)]
        print(f"{name} half={half}: {ms:.3f} ms/iter  {bpv * nv / ms / 1e6:.0f} GB/s (algorithmic {bpv} B/voxel)")
