"""k_rof_tv3d_w with packed fp32 row arithmetic (TMB_LIB = the variant build) against the shipped library: the result of 30
iterations must be bit-identical (same roundings); ms per iteration at 2048^2 x 512 and 1024^2 x 256.  Run once per
library; the second run compares with the file the first one wrote."""
import os
import sys

import torch

sys.path.insert(0, ".")
from tomobar_b200.regularisersCuPy import ROF_TV_cupy  # noqa: E402

tag = sys.argv[1]
torch.manual_seed(0)
res = {}
for shape in ((9, 21, 244), (66, 37, 364), (40, 132, 8), (64, 256, 256)):
    for lam, tau, scale in ((3e-4, 1e-3, 0.02), (0.05, 0.02, 1.0)):
        g = torch.Generator(device="cuda").manual_seed(shape[0] + shape[2])
        v = torch.randn(*shape, device="cuda", generator=g) * scale
        res[f"{shape}-{lam}"] = ROF_TV_cupy(v, lam, 30, tau, 0, False).cpu()
path = "/tmp/rof_ab.pt"
if os.path.exists(path):
    ref = torch.load(path)
    for k in res:
        print(f"{k}: bit-identical to the other library = {torch.equal(res[k], ref[k])}  max diff {float((res[k] - ref[k]).abs().max()):.3e}")
else:
    torch.save(res, path)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for nz, n in ((512, 2048), (256, 1024)):
    v = torch.randn(nz, n, n, device="cuda") * 0.02
    out = torch.empty_like(v)
    its = 20
    ROF_TV_cupy(v, 3e-4, its, 1e-3, 0, False, out=out)
    torch.cuda.synchronize()
    a.record()
    for _ in range(3):
        ROF_TV_cupy(v, 3e-4, its, 1e-3, 0, False, out=out)
    b.record()
    torch.cuda.synchronize()
    print(f"[{tag}] ROF_TV {n}^2 x {nz}: {a.elapsed_time(b) / 3 / its:.3f} ms per iteration", flush=True)
    del v, out
