"""Kernel-level reproducibility of the gather at (n na nz2): the same call repeated on the same input, per variant."""
import math
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from tomobar_b200._lib import lib, check  # noqa: E402
from tomobar_b200._tensors import ptr  # noqa: E402

n, na, nz2 = (int(v) for v in sys.argv[1:4])
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream(dev).cuda_stream
theta = torch.as_tensor(-np.linspace(0, math.pi, na, endpoint=False).astype(np.float32), dtype=torch.float32, device=dev)
sorted_theta, sorted_idx = torch.sort(theta)
sorted_idx = sorted_idx.to(torch.int32)
g = torch.Generator(device="cuda").manual_seed(1)
datac = torch.view_as_complex(torch.randn((nz2, na, n, 2), device=dev, generator=g))
dataz = torch.empty_like(datac)
check(lib.tmb_fi_scale_sign_pairs(ptr(datac), ptr(dataz), 1.0, n, na, nz2, st), "p")
mu = -np.log(1e-4) / (2 * n * n)
m = int(np.ceil(2 * n * 1 / np.pi * np.sqrt(-mu * np.log(1e-4) + (mu * n) * (mu * n) / 4)))
for name, fn, src, mode, sc in (("k_fi_gather", lib.tmb_fi_gather, datac, 1, 0), ("gather_w sc 4", lib.tmb_fi_gather, datac, 3, 4),
                                ("gather_w sc 8", lib.tmb_fi_gather, datac, 3, 8), ("gather_w sc 8 predicated", lib.tmb_fi_gather, datac, 3, 108),
                                ("gather_w default", lib.tmb_fi_gather, datac, 0, 0), ("pairs sc 8", lib.tmb_fi_gather_pairs, dataz, 0, 8),
                                ("pairs default", lib.tmb_fi_gather_pairs, dataz, 0, 0)):
    lib.tmb_fi_set_gather(mode), lib.tmb_fi_set_slices_per_thread(sc)
    outs = []
    for k in range(6):
        fde = torch.full((nz2, 2 * n, 2 * n), float("nan"), dtype=torch.complex64, device=dev)
        junk = torch.empty(4096 * (k + 1), device=dev)
        check(fn(ptr(src), ptr(fde), ptr(theta), ptr(sorted_theta), ptr(sorted_idx), m, float(np.float32(mu)), n, na, nz2, st), "g")
        outs.append(torch.view_as_real(fde))
    lib.tmb_fi_set_gather(0), lib.tmb_fi_set_slices_per_thread(0)
    same = [torch.equal(outs[0], o) for o in outs[1:]]
    nd = [int(((outs[0] - o).abs() > 0).sum()) for o in outs[1:]]
    print(f"{name:28s} runs equal to run 0: {same}  differing values: {nd}")
