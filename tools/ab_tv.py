"""A/B timing of two builds of libtmb on the same box: python tools/ab_tv.py  (spawns itself with TMB_LIB)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch

    sys.path.insert(0, HERE)
    from tomobar_b200.regularisersCuPy import PD_TV_cupy, ROF_TV_cupy

    nz, n = 512, 2048
    v = torch.randn(nz, n, n, device="cuda") * 0.02
    out = torch.empty_like(v)

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    for half in (False, True):
        ms50 = timed(lambda: PD_TV_cupy(v, 3e-4, 50, 0, 1, 12.0, 0, half, out=out))
        print(f"  PD_TV x50 half={int(half)}: {ms50:8.2f} ms per prox call ({ms50 / 50:.3f} ms/iter incl. setup)", flush=True)
    ms30 = timed(lambda: ROF_TV_cupy(v, 3e-4, 30, 1e-3, 0, False, out=out))
    print(f"  ROF_TV x30: {ms30:8.2f} ms per prox call", flush=True)
else:
    for rep in range(2):
        for lib in ("libtmb_prev.so", "libtmb.so"):
            env = dict(os.environ, TMB_LIB=os.path.join(HERE, "tomobar_b200", lib))
            print(lib, flush=True)
            subprocess.run([sys.executable, __file__, "child"], env=env, check=False)
