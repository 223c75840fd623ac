"""k_fi_gather_s (samples of a tile staged in shared memory, hook 2) against k_fi_gather (hook 1): agreement of the
grids and ms per gather at config 4 (2048^2 x 128, 2000 angles) and two smaller shapes; then the whole FOURIER_INV call."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from tomobar_b200._lib import lib, check  # noqa: E402
from tomobar_b200._tensors import ptr  # noqa: E402

dev = torch.device("cuda", 0)
st = torch.cuda.current_stream(dev).cuda_stream
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
SHAPES = ((2048, 2000, 128, np.pi), (362, 241, 10, np.pi), (256, 180, 6, 2 * np.pi), (1024, 900, 64, np.pi))
if len(sys.argv) >= 4:  # python tools/check_gather.py n nangles nz
    SHAPES = ((int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), np.pi),)
for n, na, nz, span in SHAPES:
    nz2 = nz // 2
    angles = np.linspace(0, span, na, endpoint=False).astype(np.float32)
    theta = torch.as_tensor(-angles, dtype=torch.float32, device=dev)
    sorted_theta, sorted_idx = torch.sort(theta)
    sorted_idx = sorted_idx.to(torch.int32)
    datac = torch.randn((nz2, na, n), dtype=torch.complex64, device=dev)
    mu = -np.log(1e-4) / (2 * n * n)
    m = int(np.ceil(2 * n * 1 / np.pi * np.sqrt(-mu * np.log(1e-4) + (mu * n) * (mu * n) / 4)))
    fde = torch.empty((nz2, 2 * n, 2 * n), dtype=torch.complex64, device=dev)
    ref = None
    # mode 3x / 31x: k_fi_gather_w with x complex slices per thread; 4xx: the same with the per-slice predicates kept on
    # full chunks (sc + 100 of the hook)
    for mode in (1, 2, 3, 38, 316, 408, 416, 0):
        lib.tmb_fi_set_gather(3 if mode > 3 else mode)
        sc = 0
        if mode == 3:
            sc = 4
        elif mode >= 400:
            sc = 100 + mode - 400
        elif mode > 300:
            sc = mode - 300
        elif mode > 3:
            sc = mode - 30
        lib.tmb_fi_set_slices_per_thread(sc)
        fn = lambda: check(lib.tmb_fi_gather(ptr(datac), ptr(fde), ptr(theta), ptr(sorted_theta), ptr(sorted_idx), m,
                                             float(np.float32(mu)), n, na, nz2, st), "g")
        fde.fill_(float("nan"))
        fn(); torch.cuda.synchronize()
        a.record()
        for _ in range(3):
            fn()
        b.record(); torch.cuda.synchronize()
        r = torch.view_as_real(fde)
        if ref is None:
            ref = r.clone()
        d = (r - ref)
        print(f"n={n} na={na} nz={nz} m={m} mode {mode}: {a.elapsed_time(b) / 3:8.2f} ms  rel-L2 vs mode 1 "
              f"{(d.norm() / ref.norm()).item():.2e}  max {d.abs().max().item():.2e} (|ref| max {ref.abs().max().item():.2e}) "
              f"finite={bool(torch.isfinite(r).all())}", flush=True)
    lib.tmb_fi_set_gather(0)
    if nz2 % 8 == 0:  # the slice-pair layout (tmb_fi_scale_sign_pairs -> tmb_fi_gather_pairs), 8 and 16 slices per thread
        dataz = torch.empty_like(datac)
        check(lib.tmb_fi_scale_sign_pairs(ptr(datac), ptr(dataz), 1.0, n, na, nz2, st), "pairs")
        check(lib.tmb_fi_scale_sign(ptr(datac), 1.0, n, na, nz2, st), "planar")
        lib.tmb_fi_set_gather(1)
        check(lib.tmb_fi_gather(ptr(datac), ptr(fde), ptr(theta), ptr(sorted_theta), ptr(sorted_idx), m,
                                float(np.float32(mu)), n, na, nz2, st), "g")
        ref = torch.view_as_real(fde).clone()
        lib.tmb_fi_set_gather(0)
        for sc in (8, 16):
            lib.tmb_fi_set_slices_per_thread(sc)
            fn = lambda: check(lib.tmb_fi_gather_pairs(ptr(dataz), ptr(fde), ptr(theta), ptr(sorted_theta), ptr(sorted_idx),
                                                       m, float(np.float32(mu)), n, na, nz2, st), "gp")
            fde.fill_(float("nan"))
            fn(); torch.cuda.synchronize()
            a.record()
            for _ in range(3):
                fn()
            b.record(); torch.cuda.synchronize()
            d = torch.view_as_real(fde) - ref
            print(f"n={n} na={na} nz={nz} m={m} mode pairs{sc}: {a.elapsed_time(b) / 3:8.2f} ms  rel-L2 vs mode 1 "
                  f"{(d.norm() / ref.norm()).item():.2e}  max {d.abs().max().item():.2e} (|ref| max {ref.abs().max().item():.2e}) "
                  f"finite={bool(torch.isfinite(fde.real).all())}", flush=True)
        del dataz
    lib.tmb_fi_set_slices_per_thread(0)
    del datac, fde, ref, r, d
    torch.cuda.empty_cache()
