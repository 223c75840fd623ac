"""Times the direct methods (CUDA events): FBP and FOURIER_INV at BASELINE.json's config 4 size
(2048 x 2048 x 128, 2000 angles) or `python tools/bench_direct.py n nz nangles`."""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tomobar_b200.methodsDIR_CuPy import RecToolsDIRCuPy  # noqa: E402


def timed(fn, reps=2):
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
    n, nz, na = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (2048, 128, 2000)
    angles = np.linspace(0.0, math.radians(179.9), na).astype(np.float32)
    R = RecToolsDIRCuPy(n, 0, nz, 0.0, angles, n, device_projector=0)
    g = torch.Generator(device="cuda").manual_seed(0)
    data = torch.rand((nz, na, n), device="cuda", generator=g)
    upd = float(nz) * n * n * na
    ms = timed(lambda: R.FOURIER_INV(data, filter_type="shepp", cutoff_freq=1.0))
    print(f"FOURIER_INV {n}x{n}x{nz}, {na} angles: {ms:9.2f} ms  ({nz / ms * 1e3:8.1f} slices/s)", flush=True)
    ms = timed(lambda: R.FBP(data, data_axes_labels_order=["detY", "angles", "detX"], cutoff_freq=1.0))
    print(f"FBP         {n}x{n}x{nz}, {na} angles: {ms:9.2f} ms  ({nz / ms * 1e3:8.1f} slices/s, "
          f"{upd / ms / 1e6:8.1f} GUPS incl. filter)", flush=True)
    ms = timed(lambda: R.BACKPROJ(data))
    print(f"BACKPROJ    {n}x{n}x{nz}, {na} angles: {ms:9.2f} ms  ({upd / ms / 1e6:8.1f} GUPS, "
          f"{float(nz) * na * n / ms / 1e6:6.2f} GProj/s)", flush=True)
    vol = torch.rand((nz, n, n), device="cuda", generator=g)
    from tomobar_b200._lib import lib

    for mode, name in ((0, "default"), (1, "k_fp"), (4, "k_fpq<1>"), (3, "k_fpq<2>"), (2, "k_fpq seg")):
        lib.tmb_fp_set_kernel(mode)
        try:
            Rm = RecToolsDIRCuPy(n, 0, nz, 0.0, angles, n, device_projector=0)
            ms = timed(lambda: Rm.FORWPROJ(vol))
        finally:
            lib.tmb_fp_set_kernel(0)
        print(f"FORWPROJ {name:9s} {n}x{n}x{nz}, {na} angles: {ms:9.2f} ms  ({upd / ms / 1e6:8.1f} GUPS, "
              f"{float(nz) * na * n / ms / 1e6:6.2f} GProj/s)", flush=True)
        del Rm
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
