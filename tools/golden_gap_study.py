"""Which modelling term moves the reference goldens that still miss rtol 1e-4?  (VERDICT r1, weak point 1.)

Runs the ORACLE (CPU) on the two 2-D FISTA goldens of tests/test_RecToolsIRCuPy.py:358-388, 580-611 under the
switches of oracle/proj_oracle.c and prints the distance to the golden for each:
  quant 1  the shipped model: 8-bit texture weights (what libtmb's kernels compute)
  quant 3  + ASTRA's accumulation order (32-line slab sums scaled and added in FP, 32-angle group sums in BP)
  quant 5  + interpolation coordinates rounded differently (no FMA, other association): ~0.2 % of the weights flip
  quant 7  both
  quant 0  exact fp32 weights (no texture quantisation)
  quant 41 / 273  the 8-bit weights truncated / rounded half-up instead of rounded to nearest-even
  tie      marching direction at |sin| == |cos| (angles 45 and 135 of the scan tie exactly in fp32) flipped
  seeds    power method from six different random starts
usage: python tools/golden_gap_study.py > profiles/golden_gap_study_r02.txt"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

d = np.load(os.path.join(ROOT, "tests", "golden", "normalised_data.npz"))
sino, ang = d["data_norm"][:, 64, :], d["angles"]
CASES = {"fista_2d_x50": (None, 50, -0.010516173, 0.03179016), "fista_os5_2d": (5, 10, -0.010578496, 0.03182499)}


def run(q, os_n, its, seed=0):
    rec = O.RecIR(160, 0, None, 0.0, ang, 160, os_n, quant=q)
    lc = rec.powermethod(seed=seed)
    x = rec.FISTA(sino[None], its, lipschitz_const=lc)
    return lc, float(x.min()), float(x.max())


for name, (os_n, its, gmin, gmax) in CASES.items():
    print(f"== {name}: golden min {gmin} max {gmax} (reference rtol 1e-6)")
    base = None
    for q, label in ((1, "8-bit weights (shipped model)"), (3, "+ ASTRA accumulation order"),
                     (5, "+ other fp32 rounding of the coordinates"), (7, "+ both"), (0, "exact weights"),
                     (1 | 8 | 32, "weights truncated to 8 bits, not rounded"),
                     (1 | 16 | 256, "weights rounded half-up, not half-even")):
        lc, mn, mx = run(q, os_n, its)
        base = base or mn
        print(f"  quant {q:3d} {label:42s} L {lc:.3f}  min {mn:.9f} rel {mn / gmin - 1:+.2e}  max {mx:.9f} rel {mx / gmax - 1:+.2e}"
              f"  | min moved by {mn / base - 1:+.1e} from the shipped model")
    orig = O.angle_table

    def tie_flipped(angles, cor, n, nu):
        t = orig(angles, cor, n, nu)
        a = np.asarray(angles)
        tie = np.abs(np.sin(a)) == np.abs(np.cos(a))
        if tie.any():  # rebuild those rows as x-marching ones
            sa, ca = np.sin(a).astype(np.float64), np.cos(a).astype(np.float64)
            for i in np.nonzero(tie)[0]:
                alpha = -ca[i] / sa[i]
                t[i, 3:] = [alpha, (-nu / 2.0 + 0.5 + cor) / sa[i] + (n / 2.0 - 0.5), 1.0 / sa[i], np.sqrt(1 + alpha * alpha), 0.0]
        return t

    O.angle_table = tie_flipped
    lc, mn, mx = run(1, os_n, its)
    O.angle_table = orig
    print(f"  tie rule flipped at angles 45, 135 {'':19s} L {lc:.3f}  min {mn:.9f} rel {mn / gmin - 1:+.2e}  max {mx:.9f} rel {mx / gmax - 1:+.2e}"
          f"  | min moved by {mn / base - 1:+.1e}")
    mins = [run(1, os_n, its, seed)[1] for seed in range(6)]
    print(f"  power method, seeds 0..5: min spread {max(mins) - min(mins):.1e} (L converges to the last bit in 15 iterations)")
print("""
Reading: none of the candidate terms moves the minimum by more than 6e-6 relative, the golden is 4.2e-4 away; removing the
8-bit weight quantisation moves it by 6.5e-4 in the WRONG direction and truncating the weights by 1.5e-3 (so the rounded
8-bit model is the right one; half-up instead of half-even rounding is worth 6e-5).  The gap is
therefore not ASTRA's accumulation order (SURVEY.md section 7, hard part 1), not coordinate rounding, not the marching-
direction tie and not the power method.  What remains is something in ASTRA's own kernels that no reference golden pins
(no golden exists for a forward projection of a non-constant volume): both cases are the converged least-squares solutions
of a single slice, where a static operator difference of ~1e-6 is amplified at the noisiest pixel (min at (47, 106), 42
pixels from the centre; the maximum agrees to 1e-5).  The CUDA kernels reproduce the oracle to <= 2e-6 on these cases.""")
