"""k_fpm segment-length / group-size sweep at the headline and config-2 sizes (ms per subset forward projection)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from tomobar_b200._lib import lib  # noqa: E402
from tomobar_b200.projector import ProjTools3D  # noqa: E402

for nz, n, na, os_n, segs in ((512, 2048, 1800, 24, (0, 216, 162, 108)), (256, 1024, 900, 6, (0, 304, 204, 152))):
    angles = np.linspace(0, np.pi, na, endpoint=False).astype(np.float32)
    vol = torch.rand((nz, n, n), device="cuda")
    for seg in segs:
        for mode in (2, 6, 7):
            lib.tmb_fp_set_kernel(mode)
            lib.tmb_fp_set_segment(seg)
            try:
                P = ProjTools3D(n, 0, nz, angles, 0.0, n, "gpu", 0, os_n)
                grp = lib.tmb_geom_fp_group(P._g, 1)
                out = P._forwprojOSCuPy(vol, 1)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(5):
                    out = P._forwprojOSCuPy(vol, 1)
                b.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b) / 5
            finally:
                lib.tmb_fp_set_kernel(0)
                lib.tmb_fp_set_segment(0)
            upd = float(nz) * n * n * out.shape[1]
            print(f"FP mode {mode} group={grp} seg={seg} {n}x{n}x{nz} {out.shape[1]} angles: {ms:8.2f} ms  "
                  f"{upd / ms / 1e9:6.3f} TUPS", flush=True)
            del P, out
            torch.cuda.empty_cache()
    del vol
    torch.cuda.empty_cache()
