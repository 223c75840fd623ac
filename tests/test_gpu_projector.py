"""Parity of the CUDA projector pair (through the C ABI) with the CPU oracle on seeded inputs."""

import numpy as np
import pytest
import torch

from conftest import rel_l2, rel_max

pytestmark = pytest.mark.gpu

# float tolerance of BASELINE.json's north star: 1e-4 relative; the kernels share the oracle's
# operation order, so they are held to a much tighter bound here.
TOL = 2e-6

CASES = [
    # nz, n, nu, na, cor, os
    (1, 32, 32, 16, 0.0, None),
    (3, 40, 56, 30, 0.0, None),
    (8, 64, 64, 45, 1.5, None),
    (13, 75, 91, 61, -2.25, None),
    (16, 160, 160, 180, 0.0, None),
    (9, 50, 70, 37, 0.5, 5),
    (8, 200, 260, 90, 3.0, 6),
    (33, 70, 100, 40, 0.75, 4),   # > 16 slices: the 32-slice-blocked forward projector (k_fpq)
    (40, 96, 64, 24, 0.0, None),
]


def _angles(na, full=False):
    end = 2 * np.pi if full else np.pi
    return np.linspace(0, end, na, endpoint=False).astype(np.float32)


@pytest.mark.parametrize("nz,n,nu,na,cor,os_n", CASES)
def test_fp_bp_match_oracle(oracle, nz, n, nu, na, cor, os_n):
    from tomobar_b200.projector import ProjTools3D

    rng = np.random.default_rng(nz * 1000 + n)
    angles = _angles(na, full=(n == 75))
    P = ProjTools3D(nu, 0, nz, angles, cor, n, "gpu", 0, os_n)
    O = oracle.Atools(nu, 0, nz, angles, cor, n, os_n)
    vol = rng.standard_normal((nz, n, n)).astype(np.float32)
    subsets = [None] if os_n is None else list(range(os_n))
    for s in subsets:
        if s is None:
            fp_ref = O._forwprojCuPy(vol)
            fp = P._forwprojCuPy(torch.from_numpy(vol).cuda()).cpu().numpy()
        else:
            fp_ref = O._forwprojOSCuPy(vol, s)
            fp = P._forwprojOSCuPy(torch.from_numpy(vol).cuda(), s).cpu().numpy()
        assert fp.shape == fp_ref.shape
        assert rel_max(fp, fp_ref) < TOL, f"FP subset {s}"
        sino = rng.standard_normal(fp_ref.shape).astype(np.float32)
        if s is None:
            bp_ref = O._backprojCuPy(sino)
            bp = P._backprojCuPy(torch.from_numpy(sino).cuda()).cpu().numpy()
        else:
            bp_ref = O._backprojOSCuPy(sino, s)
            bp = P._backprojOSCuPy(torch.from_numpy(sino).cuda(), s).cpu().numpy()
        assert bp.shape == (nz, n, n)
        assert rel_max(bp, bp_ref) < TOL, f"BP subset {s}"


def test_exact_weights_mode(oracle):
    from tomobar_b200.projector import ProjTools3D

    rng = np.random.default_rng(5)
    nz, n, nu, na = 4, 48, 48, 33
    angles = _angles(na)
    P = ProjTools3D(nu, 0, nz, angles, 0.0, n, "gpu", 0, None, quantise_weights=False)
    O = oracle.Atools(nu, 0, nz, angles, 0.0, n, None, quant=False)
    vol = rng.standard_normal((nz, n, n)).astype(np.float32)
    assert rel_max(P._forwprojCuPy(torch.from_numpy(vol).cuda()).cpu().numpy(), O._forwprojCuPy(vol)) < TOL


def test_fused_gradient_matches_unfused(oracle):
    from tomobar_b200.projector import ProjTools3D

    rng = np.random.default_rng(11)
    nz, n, nu, na, os_n = 6, 64, 80, 40, 4
    angles = _angles(na)
    P = ProjTools3D(nu, 0, nz, angles, 0.75, n, "gpu", 0, os_n)
    x = torch.from_numpy(rng.standard_normal((nz, n, n)).astype(np.float32)).cuda()
    b = torch.from_numpy(rng.standard_normal((nz, na, nu)).astype(np.float32)).cuda()
    w = torch.rand((nz, na, nu), device="cuda")
    for s in range(os_n):
        ind = torch.arange(s, na, os_n, device="cuda")
        res = P._forwprojOSCuPy(x, s) - b[:, ind, :]
        g_ref = P._backprojOSCuPy(res, s)
        g = P.grad_data_term(x, b, s, "LS")
        assert torch.equal(g, g_ref)
        g_ref = P._backprojOSCuPy(res * w[:, ind, :], s)
        g = P.grad_data_term(x, b, s, "PWLS", w)
        assert torch.equal(g, g_ref)
    # KL
    xp = x.abs() + 0.1
    bp = b.abs()
    ind = torch.arange(1, na, os_n, device="cuda")
    res = 1 - bp[:, ind, :] / torch.clamp(P._forwprojOSCuPy(xp, 1), min=1e-8)
    g_ref = P._backprojOSCuPy(res, 1)
    g = P.grad_data_term(xp, bp, 1, "KL")
    assert rel_max(g.cpu().numpy(), g_ref.cpu().numpy()) < 1e-5


def test_inputs_untouched_and_fresh_outputs():
    from tomobar_b200.projector import ProjTools3D

    angles = _angles(20)
    P = ProjTools3D(32, 0, 4, angles, 0.0, 32)
    v = torch.rand(4, 32, 32, device="cuda")
    v0 = v.clone()
    a = P._forwprojCuPy(v)
    b = P._forwprojCuPy(v)
    assert torch.equal(v, v0) and a.data_ptr() != b.data_ptr() and torch.equal(a, b)
    # strided (swapaxes view) input is treated as the logical array
    s = torch.rand(20, 4, 32, device="cuda")
    r1 = P._backprojCuPy(s.swapaxes(0, 1))
    r2 = P._backprojCuPy(s.swapaxes(0, 1).contiguous())
    assert torch.equal(r1, r2)
    with pytest.raises(ValueError):
        P._backprojCuPy(s)
    with pytest.raises(ValueError):
        P._forwprojCuPy(v.double())


def test_linearity_and_zero_at_scale():
    """Size-independent properties on a larger problem than the oracle comfortably handles."""
    from tomobar_b200.projector import ProjTools3D

    nz, n, na = 24, 512, 360
    angles = _angles(na)
    P = ProjTools3D(n, 0, nz, angles, 0.0, n, "gpu", 0, 6)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(nz, n, n, device="cuda", generator=g)
    y = torch.randn(nz, n, n, device="cuda", generator=g)
    for s in (0, 5):
        fx, fy = P._forwprojOSCuPy(x, s), P._forwprojOSCuPy(y, s)
        fxy = P._forwprojOSCuPy(2.0 * x - 3.0 * y, s)
        assert rel_max((2.0 * fx - 3.0 * fy).cpu().numpy(), fxy.cpu().numpy()) < 1e-5
        assert torch.count_nonzero(P._forwprojOSCuPy(torch.zeros_like(x), s)) == 0
        bx = P._backprojOSCuPy(fx, s)
        assert torch.isfinite(bx).all()
    # slices are independent: permuting the stack permutes the projection (bit for bit), and projecting
    # a sub-stack equals slicing the projection (an 8-slice stack runs k_fp, the 24-slice one the
    # line-segmented k_fpq, which adds its partial sums in a different order: fp32 rounding)
    assert torch.equal(P._forwprojOSCuPy(x.flip(0).contiguous(), 3), P._forwprojOSCuPy(x, 3).flip(0))
    P2 = ProjTools3D(n, 0, 8, angles, 0.0, n, "gpu", 0, 6)
    assert rel_max(P2._forwprojOSCuPy(x[8:16].contiguous(), 3).cpu().numpy(),
                   P._forwprojOSCuPy(x, 3)[8:16].cpu().numpy()) < 2e-6
    # unmatched pair is still close to adjoint: <Ax, y> ~ <x, A^T y>
    full = ProjTools3D(n, 0, nz, angles, 0.0, n)
    # (smooth inputs: on white noise the Joseph / voxel-driven pair differs by >10 %)
    import torch.nn.functional as F
    xs = F.avg_pool2d(x[None], 9, 1, 4)[0].contiguous()
    ax = full._forwprojCuPy(xs)
    q = F.avg_pool2d(torch.randn(ax.shape, device="cuda", generator=g)[None], (1, 9), 1, (0, 4))[0].contiguous()
    x = xs
    lhs = torch.sum(ax.double() * q.double()).item()
    rhs = torch.sum(x.double() * full._backprojCuPy(q).double()).item()
    assert abs(lhs - rhs) / max(abs(lhs), abs(rhs)) < 5e-2


@pytest.mark.parametrize("nz,n,nu,na,os_n", [(5, 64, 80, 36, 3), (20, 130, 130, 50, None), (70, 48, 40, 21, 2),
                                             (64, 40, 64, 12, None)])
def test_forward_projector_kernels_agree(nz, n, nu, na, os_n):
    """k_fp (8 slices per thread) and k_fpq (bank-conflict-free 32-slice blocks): one block per CTA
    (mode 4), two (mode 3) are bit-identical to k_fp; the line-segmented k_fpq (mode 2, segments
    forced short here) adds its partial sums in a different order and agrees to fp32 rounding; the
    multi-angle k_fpm (modes 5-7, the default where its windows fit) is bit-identical to mode 2."""
    from tomobar_b200._lib import lib
    from tomobar_b200.projector import ProjTools3D

    g = torch.Generator(device="cuda").manual_seed(nz)
    vol = torch.randn((nz, n, n), device="cuda", generator=g)
    b = torch.randn((nz, na, nu), device="cuda", generator=g)
    w = torch.rand((nz, na, nu), device="cuda", generator=g)
    res = {}
    for mode in (1, 2, 3, 4, 5, 6, 7):
        lib.tmb_fp_set_kernel(mode)
        lib.tmb_fp_set_segment(24 if mode in (2, 5, 6, 7) else 0)
        try:
            P = ProjTools3D(nu, 0, nz, _angles(na), 0.5, n, "gpu", 0, os_n)
            sub = None if os_n is None else os_n - 1
            fp = P._forwprojCuPy(vol) if os_n is None else P._forwprojOSCuPy(vol, sub)
            res[mode] = (fp, P.grad_data_term(vol, b, sub, "PWLS", w), P.grad_data_term(vol, b.abs(), sub, "KL"))
        finally:
            lib.tmb_fp_set_kernel(0)
            lib.tmb_fp_set_segment(0)
    for mode in (3, 4):
        for a, c in zip(res[1], res[mode]):
            assert torch.equal(a, c)
    for a, c in zip(res[1][:2], res[2][:2]):
        assert rel_max(c.cpu().numpy(), a.cpu().numpy()) < 2e-6
    # k_fpm (2 / 3 / 4 angles of the subset per CTA sharing one window): same arithmetic and summation order as
    # the segmented k_fpq
    for mode in (5, 6, 7):
        for a, c in zip(res[2], res[mode]):
            assert torch.equal(a, c)
