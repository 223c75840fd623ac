"""ncu target: a few PD_TV (fp32 / fp16 duals) and ROF_TV iterations at a given size (default: the
headline 512 x 2048 x 2048).  python tools/prof_tv_big.py [nz n]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tomobar_b200.regularisersCuPy import PD_TV_cupy, ROF_TV_cupy  # noqa: E402

nz, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (512, 2048)
x = torch.rand(nz, n, n, device="cuda") * 0.02
out = torch.empty_like(x)
PD_TV_cupy(x, 3e-4, 3, 0, 1, 12.0, 0, False, out=out)
PD_TV_cupy(x, 3e-4, 3, 0, 1, 12.0, 0, True, out=out)
ROF_TV_cupy(x, 3e-4, 3, 1e-3, 0, False, out=out)
torch.cuda.synchronize()
print("done")
