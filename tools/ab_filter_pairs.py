"""FOURIER_INV at BASELINE.json's config 4 (2048 x 2048 x 128, 2000 angles) with STEP 0 on complex slice-pair rows
(default) and on an rfft / irfft pair per slice; CUDA events, whole call and the filter stage alone."""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tomobar_b200.methodsDIR_CuPy import RecToolsDIRCuPy  # noqa: E402


def timed(fn, reps=5):
    fn()
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
    datac = torch.empty((nz // 2, na, n), dtype=torch.complex64, device="cuda")
    res = {}
    for pairs in (True, False, True, False):
        R._FILTER_SLICE_PAIRS = pairs
        f = timed(lambda: R._fourier_filter(data, n, n, True, 4, "shepp", 1.0, pack_into=datac))
        torch.cuda.reset_peak_memory_stats()
        t = timed(lambda: R.FOURIER_INV(data, filter_type="shepp", cutoff_freq=1.0))
        res[pairs] = R.FOURIER_INV(data)
        print(f"slice pairs {pairs!s:5}: filter stage {f:7.2f} ms, FOURIER_INV {t:7.2f} ms ({nz / t * 1e3:7.1f} slices/s), "
              f"peak {torch.cuda.max_memory_allocated() / 1e9:.3f} GB", flush=True)
    a, b = res[True], res[False]
    print(f"rel-L2 between the two: {float((a - b).norm() / b.norm()):.3e}, max {float((a - b).abs().max() / b.abs().max()):.3e}")


if __name__ == "__main__":
    main()
