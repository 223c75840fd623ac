"""Per-call timing spread of the forward projector variants (min / median / max over repetitions)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from tomobar_b200._lib import lib  # noqa: E402
from tomobar_b200.projector import ProjTools3D  # noqa: E402

cfgs = ((512, 2048, 1800, 24), (256, 1024, 900, 6), (64, 2048, 1800, 24))
reps = 12
for nz, n, na, os_n in cfgs:
    angles = np.linspace(0, np.pi, na, endpoint=False).astype(np.float32)
    vol = torch.rand((nz, n, n), device="cuda")
    for rnd in range(2):
        for mode in (2, 5, 6, 7):
            lib.tmb_fp_set_kernel(mode)
            try:
                P = ProjTools3D(n, 0, nz, angles, 0.0, n, "gpu", 0, os_n)
                out = P._forwprojOSCuPy(vol, 1)
                torch.cuda.synchronize()
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
                ev[0].record()
                for r in range(reps):
                    out = P._forwprojOSCuPy(vol, (1 + r) % os_n)
                    ev[r + 1].record()
                torch.cuda.synchronize()
                ms = sorted(ev[r].elapsed_time(ev[r + 1]) for r in range(reps))
            finally:
                lib.tmb_fp_set_kernel(0)
            upd = float(nz) * n * n * out.shape[1]
            print(f"FP mode {mode} {n}x{n}x{nz} {out.shape[1]} angles: min {ms[0]:7.2f} med {ms[reps // 2]:7.2f} max {ms[-1]:7.2f} ms  "
                  f"{upd / ms[reps // 2] / 1e9:6.3f} TUPS (median)", flush=True)
            del P, out
            torch.cuda.empty_cache()
    del vol
    torch.cuda.empty_cache()
