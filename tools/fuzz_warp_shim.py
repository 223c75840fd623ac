"""Random shapes / z-runs / flags through the fused PD_TV kernels' CUDA source on the CPU warp shim
(tests/warp_shim, built by tests/test_warp_shim_fused_tv.py) against two plain iterations.

    python tools/fuzz_warp_shim.py whole   <seed> <seconds>     # variants 0, 1, 2, 4 on whole volumes
    python tools/fuzz_warp_shim.py sharded <seed> <seconds>     # the GHOST variant on 2-3 z-shards

Round 1: 1173 whole-volume and 482 sharded cases, no mismatch."""
import ctypes as C
import importlib.util
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_warp_shim_fused_tv as T  # noqa: E402

lib = C.CDLL(os.path.join(ROOT, "tests", "warp_shim", "_build", "libshim_fused_tv.so"))
lib.shim_run_fused_tv.restype = C.c_int
lib.shim_run_fused_tv.argtypes = ([C.c_int] * 3 + [T.FP] * 9 + [C.c_float] * 4 + [C.c_int] * 6 + [T.FP] * 10)
spec = importlib.util.spec_from_file_location("emu", os.path.join(ROOT, "tools", "emulate_pd_fused2.py"))
emu = importlib.util.module_from_spec(spec)
spec.loader.exec_module(emu)


def close(a, b):
    return bool(np.isfinite(a).all() and np.max(np.abs(a - b)) <= 2e-6 * max(np.max(np.abs(b)), 1.0))


def run(variant, nonneg, aniso, s, zrun, lo, hi, ghost):
    dz, dy, dx = s["U"].shape
    Uo = T._aligned(s["U"].shape, np.nan)
    Q = [T._aligned(s["U"].shape, np.nan) for _ in range(3)]
    lib.shim_run_fused_tv(variant, int(nonneg), int(aniso), T._ptr(s["inp"]), T._ptr(s["U"]), T._ptr(Uo),
                          *[T._ptr(p) for p in s["P"]], *[T._ptr(q) for q in Q], T.SIGMA, T.TAU, T.LT, T.THETA,
                          dx, dy, dz, zrun, int(lo), int(hi), *ghost)
    return Uo, Q


def main():
    mode, seed, seconds = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
    rng = np.random.default_rng(seed)
    t0, n, bad = time.time(), 0, 0
    while time.time() - t0 < seconds:
        nonneg, aniso = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        dy, dx = int(rng.integers(2, 38)), 4 * int(rng.integers(1, 66))
        if mode == "whole":
            dz = int(rng.integers(2, 11))
            sizes, variant = [dz], int(rng.choice([0, 1, 2, 4]))
        else:
            sizes, variant = [int(rng.integers(2, 6)) for _ in range(int(rng.integers(2, 4)))], 3
            dz = sum(sizes)
        zrun = int(rng.integers(1, max(sizes) + 1))
        shape = (dz, dy, dx)
        inp, U, P = T._case(shape, int(rng.integers(0, 1 << 30)))
        U2, P2 = T._two_plain(emu, inp, U, P, nonneg, aniso)
        edges = np.concatenate([[0], np.cumsum(sizes)]).astype(int)
        S = [dict(inp=T._aligned((b - a, dy, dx), inp[a:b]), U=T._aligned((b - a, dy, dx), U[a:b]),
                  P=[T._aligned((b - a, dy, dx), p[a:b]) for p in P], n=int(b - a))
             for a, b in zip(edges[:-1], edges[1:])]
        outs = []
        for i, s in enumerate(S):
            lo = S[i - 1] if i > 0 else None
            hi = S[i + 1] if i + 1 < len(S) else None
            g = [None] * 10
            if lo is not None:
                g[0] = T._ptr(lo["U"][lo["n"] - 2:])
                g[1:4] = [T._ptr(p[lo["n"] - 2:]) for p in lo["P"]]
                g[4] = T._ptr(lo["inp"][lo["n"] - 1:])
            if hi is not None:
                g[5] = T._ptr(hi["U"])
                g[6:9] = [T._ptr(p) for p in hi["P"]]
                g[9] = T._ptr(hi["inp"])
            outs.append(run(variant, nonneg, aniso, s, zrun, lo is not None, hi is not None, g))
        ok = close(np.concatenate([o[0] for o in outs]), U2)
        ok = ok and all(close(np.concatenate([o[1][c] for o in outs]), P2[c]) for c in range(3))
        n += 1
        if not ok:
            bad += 1
            print("MISMATCH", mode, shape, sizes, zrun, nonneg, aniso, variant, flush=True)
    print(f"{mode}: {n} cases, {bad} mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
