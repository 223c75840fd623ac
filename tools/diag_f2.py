"""Time selected tmb_tv_set_simple_kernels modes of PD_TV: python tools/diag_f2.py nz n mode[:name] ...  (measurement
only; modes >= 14 are the DIAG variants whose results are garbage)"""
import sys

import torch

sys.path.insert(0, ".")
from tomobar_b200._lib import lib  # noqa: E402
from tomobar_b200.regularisersCuPy import PD_TV_cupy  # noqa: E402

nz, n = int(sys.argv[1]), int(sys.argv[2])
its = 20
v = torch.randn(nz, n, n, device="cuda") * 0.02
out = torch.empty_like(v)
for spec in sys.argv[3:]:
    mode, _, name = spec.partition(":")
    old = lib.tmb_tv_set_simple_kernels(int(mode))
    try:
        PD_TV_cupy(v, 3e-4, its, 0, 1, 12.0, 0, False, out=out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            PD_TV_cupy(v, 3e-4, its, 0, 1, 12.0, 0, False, out=out)
        b.record()
        torch.cuda.synchronize()
    finally:
        lib.tmb_tv_set_simple_kernels(old)
    ms = a.elapsed_time(b) / 3 / its
    print(f"PD_TV mode {mode:>2s} {name:24s} {nz}x{n}x{n}: {ms:8.3f} ms/iter  {36 * v.numel() / ms / 1e6:8.1f} GB/s per-iteration-equivalent",
          flush=True)
