"""k_fpm (several angles of a subset per CTA sharing one staged window) against k_fpq: bit-equality on a set of
shapes (short forced segments, ordered subsets, CoR, 360-degree scans), then ms per subset forward projection at
the given sizes.
usage: python tools/check_fpm.py [time]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from tomobar_b200._lib import lib  # noqa: E402
from tomobar_b200.projector import ProjTools3D  # noqa: E402


def fp(mode, seg, nz, n, nu, angles, cor, os_n, sub, vol, quant=True):
    lib.tmb_fp_set_kernel(mode)
    lib.tmb_fp_set_segment(seg)
    try:
        P = ProjTools3D(nu, 0, nz, angles, cor, n, "gpu", 0, os_n, quantise_weights=quant)
        grp = lib.tmb_geom_fp_group(P._g, -1 if os_n is None else sub)
        out = P._forwprojCuPy(vol) if os_n is None else P._forwprojOSCuPy(vol, sub)
        return out, grp
    finally:
        lib.tmb_fp_set_kernel(0)
        lib.tmb_fp_set_segment(0)


def main():
    torch.manual_seed(0)
    bad = 0
    cases = [  # nz, n, nu, na, span, cor, os, subset, segment
        (33, 128, 128, 180, np.pi, 0.0, None, 0, 24),
        (40, 200, 232, 360, 2 * np.pi, 3.5, 4, 1, 48),
        (64, 96, 80, 90, np.pi, -2.0, 2, 1, 24),
        (20, 256, 256, 720, np.pi, 0.5, 6, 5, 60),
        (32, 130, 190, 400, np.pi, 0.0, 3, 2, 33),
        (70, 64, 64, 64, np.pi, 0.0, None, 0, 24),
    ]
    for nz, n, nu, na, span, cor, os_n, sub, seg in cases:
        angles = np.linspace(0, span, na, endpoint=False).astype(np.float32)
        vol = torch.randn((nz, n, n), device="cuda")
        for quant in (True, False):
            ref, _ = fp(2, seg, nz, n, nu, angles, cor, os_n, sub, vol, quant)
            for mode in (5, 6, 7):
                out, grp = fp(mode, seg, nz, n, nu, angles, cor, os_n, sub, vol, quant)
                eq = torch.equal(out, ref)
                d = (out - ref).abs().max().item() / ref.abs().max().item()
                bad += (not eq)
                print(f"mode {mode} group={grp} quant={quant} nz={nz} n={n} nu={nu} na={na} os={os_n} seg={seg}: "
                      f"bit-equal={eq} rel {d:.2e}", flush=True)
    print("MISMATCHES", bad, flush=True)
    if len(sys.argv) > 1:
        for nz, n, na, os_n in ((512, 2048, 1800, 24), (256, 1024, 900, 6), (64, 2048, 1800, 24)):
            angles = np.linspace(0, np.pi, na, endpoint=False).astype(np.float32)
            vol = torch.rand((nz, n, n), device="cuda")
            ref = None
            for mode in (2, 5, 6, 7):
                lib.tmb_fp_set_kernel(mode)
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
                if ref is None:
                    ref = out.clone()
                upd = float(nz) * n * n * out.shape[1]
                print(f"FP mode {mode} group={grp} {n}x{n}x{nz} {out.shape[1]} angles: {ms:8.2f} ms  {upd / ms / 1e9:6.3f} TUPS "
                      f"bit-equal={torch.equal(out, ref)}", flush=True)
                del P, out
                torch.cuda.empty_cache()
            del vol, ref
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
