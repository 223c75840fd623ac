"""Random geometries through k_fpm (default) and k_fpq (hook 2): the sinograms must be bit-identical whatever group
size the host picks (ordered subsets, per-angle CoR, 360-degree and irregular angle sets, short forced segments).
usage: python tools/fuzz_fpm.py [cases] [seed] [coarse]   (a third argument: few, widely spaced angles -> groups of 2 / 3 and fall-backs)"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from tomobar_b200._lib import lib  # noqa: E402
from tomobar_b200.projector import ProjTools3D  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
hist = {}
for c in range(cases):
    n = int(rng.integers(40, 300))
    nu = int(n * rng.uniform(0.8, 1.6))
    nz = int(rng.choice([17, 20, 33, 40, 64, 70]))
    na = int(rng.integers(24, 720)) if len(sys.argv) < 4 else int(rng.integers(10, 64))
    span = float(rng.choice([np.pi, 2 * np.pi, 0.6 * np.pi]))
    kind = rng.integers(0, 4)
    angles = np.linspace(0, span, na, endpoint=False)
    if kind == 1:
        angles = np.sort(rng.uniform(0, span, na))        # irregular spacing
    elif kind == 2:
        angles = angles[::-1].copy()                       # descending
    elif kind == 3:
        angles = angles + rng.uniform(-0.3, 0.3)           # offset start
    angles = angles.astype(np.float32)
    cor = float(rng.uniform(-4, 4)) if rng.random() < 0.6 else rng.uniform(-3, 3, na)
    os_n = (int(rng.choice([1, 1, 2, 3, 5, 8, 12])) if na >= 48 else 1) if len(sys.argv) < 4 else int(rng.choice([1, 2, 3]))
    sub = int(rng.integers(0, os_n))
    seg = int(rng.choice([24, 33, 48, 60, 96]))
    quant = bool(rng.integers(0, 2))
    vol = torch.randn((nz, n, n), device="cuda")
    out = {}
    for mode in (2, 0):
        lib.tmb_fp_set_kernel(mode if mode else 7)  # 7: k_fpm with groups of up to 4 (what the default picks), fp_q forced
        lib.tmb_fp_set_segment(seg)
        try:
            P = ProjTools3D(nu, 0, nz, angles, cor, n, "gpu", 0, os_n if os_n > 1 else None, quantise_weights=quant)
            grp = lib.tmb_geom_fp_group(P._g, sub if os_n > 1 else -1)
            out[mode] = (P._forwprojOSCuPy(vol, sub) if os_n > 1 else P._forwprojCuPy(vol)).clone()
        finally:
            lib.tmb_fp_set_kernel(0)
            lib.tmb_fp_set_segment(0)
        if mode == 0:
            hist[grp] = hist.get(grp, 0) + 1
    eq = torch.equal(out[0], out[2])
    bad += not eq
    if not eq:
        d = (out[0] - out[2]).abs().max().item()
        print(f"MISMATCH case {c}: n={n} nu={nu} nz={nz} na={na} span={span:.2f} kind={kind} os={os_n} sub={sub} seg={seg} "
              f"quant={quant} group={grp} max diff {d:.3e}", flush=True)
print(f"{cases} cases, {bad} mismatches, group sizes used: {dict(sorted(hist.items()))}")
