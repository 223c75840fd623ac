"""Parity of the TV proximal kernels with the numpy oracle (literal restatement of the
reference's .cu kernels)."""

import os

import numpy as np
import pytest
import torch

from conftest import rel_max

pytestmark = pytest.mark.gpu


def _vol(shape, seed=0):
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(shape).astype(np.float32)
    # piecewise-constant structure + noise, like a reconstruction
    v += np.where(rng.random(shape) > 0.5, 1.0, 0.0).astype(np.float32)
    return v * np.float32(0.02)


@pytest.mark.parametrize("shape", [(12, 33, 47), (1, 40, 52), (40, 52), (5, 1, 64), (3, 130, 129), (6, 19, 136), (2, 2, 4)])
@pytest.mark.parametrize("methodTV", [0, 1])
@pytest.mark.parametrize("nonneg", [0, 1])
@pytest.mark.parametrize("half", [False, True])
def test_pd_tv(oracle, shape, methodTV, nonneg, half):
    from tomobar_b200.regularisersCuPy import PD_TV_cupy

    v = _vol(shape, 1)
    ref = oracle.pd_tv(v, 5e-4, 12, methodTV, nonneg, 12.0, half)
    out = PD_TV_cupy(torch.from_numpy(v).cuda(), 5e-4, 12, methodTV, nonneg, 12.0, 0, half).cpu().numpy()
    assert out.shape == ref.shape
    tol = 2e-3 if half else 1e-5
    assert rel_max(out, ref) < tol


@pytest.mark.parametrize("shape", [(12, 33, 47), (1, 40, 52), (40, 52), (3, 130, 129)])
@pytest.mark.parametrize("half", [False, True])
def test_rof_tv(oracle, shape, half):
    from tomobar_b200.regularisersCuPy import ROF_TV_cupy

    v = _vol(shape, 2)
    ref = oracle.rof_tv(v, 3e-4, 15, 1e-3, half)
    out = ROF_TV_cupy(torch.from_numpy(v).cuda(), 3e-4, 15, 1e-3, 0, half).cpu().numpy()
    assert out.shape == ref.shape
    tol = 2e-3 if half else 1e-5
    assert rel_max(out, ref) < tol


def test_tv_errors_and_edge_cases():
    from tomobar_b200.regularisersCuPy import PD_TV_cupy, ROF_TV_cupy

    x = torch.rand(4, 8, 8, device="cuda")
    with pytest.raises(ValueError):
        PD_TV_cupy(x.double())
    with pytest.raises(ValueError):
        ROF_TV_cupy(x, gpu_id=-1)
    with pytest.raises(ValueError):
        PD_TV_cupy(torch.rand(2, 2, 2, 2, device="cuda"))
    # zero iterations returns the input
    assert torch.equal(PD_TV_cupy(x, 1e-3, 0), x)
    assert torch.equal(ROF_TV_cupy(x, 1e-3, 0), x)
    # input is not modified, odd/even iteration counts both land in the output
    x0 = x.clone()
    a = PD_TV_cupy(x, 1e-3, 3)
    b = PD_TV_cupy(x, 1e-3, 4)
    assert torch.equal(x, x0) and not torch.equal(a, b)


@pytest.mark.parametrize("shape", [(70, 100, 150), (33, 9, 65), (64, 64, 64), (97, 35, 260), (130, 5, 128), (40, 64, 8)])
@pytest.mark.parametrize("half", [False, True])
def test_marching_kernels_match_simple_kernels(shape, half):
    """The warp-strip kernels (dx % 4 == 0; register-fed = mode 3, TMA-fed = mode 4), the CTA-tiled
    z-marching kernels (mode 2) and the one-thread-per-voxel kernels (mode 1) share their arithmetic."""
    from tomobar_b200._lib import lib
    from tomobar_b200.regularisersCuPy import PD_TV_cupy, ROF_TV_cupy

    v = torch.from_numpy(_vol(shape, 7)).cuda()
    res = {}
    for mode in (0, 1, 2, 3, 4, 5):
        old = lib.tmb_tv_set_simple_kernels(mode)
        try:
            res[mode] = (PD_TV_cupy(v, 5e-4, 9, 0, 1, 12.0, 0, half).cpu().numpy(),
                         PD_TV_cupy(v, 5e-4, 4, 1, 0, 12.0, 0, half).cpu().numpy(),
                         ROF_TV_cupy(v, 3e-4, 9, 1e-3, 0, half).cpu().numpy())
        finally:
            lib.tmb_tv_set_simple_kernels(old)
    tol = 2e-3 if half else 2e-6
    for mode in (0, 2, 3, 4, 5):
        for a, b in zip(res[mode], res[1]):
            assert rel_max(a, b) < tol


@pytest.mark.parametrize("shape", [(2, 2, 4), (5, 9, 124), (7, 18, 132), (9, 21, 244), (6, 16, 120), (66, 37, 364),
                                   (40, 130, 8), (3, 2, 12)])
@pytest.mark.parametrize("methodTV,nonneg", [(0, 0), (0, 1), (1, 1)])
def test_fused_pairs_of_pd_iterations(shape, methodTV, nonneg):
    """k_pd_tv3d_f2 (mode 5: two iterations per pass, nothing stored in between) against single
    iterations of the strip kernel (mode 3) and of the one-thread-per-voxel kernel (mode 1), for even
    and odd iteration counts; window / strip / z-run edges at every shape."""
    from tomobar_b200._lib import lib
    from tomobar_b200.regularisersCuPy import PD_TV_cupy

    v = torch.from_numpy(_vol(shape, 11)).cuda()
    res = {}
    for mode in (1, 3, 5):
        old = lib.tmb_tv_set_simple_kernels(mode)
        try:
            res[mode] = [PD_TV_cupy(v, 5e-4, its, methodTV, nonneg, 12.0, 0, False).cpu().numpy() for its in (2, 7, 12)]
        finally:
            lib.tmb_tv_set_simple_kernels(old)
    for a, b, c in zip(res[5], res[3], res[1]):
        assert np.isfinite(a).all()
        assert rel_max(a, b) < 2e-6 and rel_max(a, c) < 2e-6


@pytest.mark.parametrize("shape", [(9, 21, 244), (66, 37, 364), (130, 64, 128), (5, 9, 124), (40, 130, 8)])
def test_split_variant_of_the_fused_kernel(shape):
    """k_pd_tv3d_f2s (mode 6, opt-in: the same two-iteration kernel with its warm-up / march / tail steps
    specialised at compile time, 7 % faster at the headline size) on the shapes it was first run on."""
    from tomobar_b200._lib import lib
    from tomobar_b200.regularisersCuPy import PD_TV_cupy

    v = torch.from_numpy(_vol(shape, 13)).cuda()
    res = {}
    for mode in (3, 6):
        old = lib.tmb_tv_set_simple_kernels(mode)
        try:
            res[mode] = [PD_TV_cupy(v, 5e-4, its, m, nn, 12.0, 0, False).cpu().numpy()
                         for its, nn, m in ((2, 1, 0), (7, 0, 0), (4, 1, 1))]
        finally:
            lib.tmb_tv_set_simple_kernels(old)
    for a, b in zip(res[6], res[3]):
        assert np.isfinite(a).all() and rel_max(a, b) < 2e-6


@pytest.mark.parametrize("mode", [0, 5, 7, 8, 9, 10, 11, 12])
@pytest.mark.parametrize("shape", [(9, 21, 244), (66, 37, 364), (2, 2, 4), (7, 18, 132)])
def test_untimed_variants_of_the_fused_kernel(shape, mode):
    from tomobar_b200._lib import lib
    from tomobar_b200.regularisersCuPy import PD_TV_cupy

    v = torch.from_numpy(_vol(shape, 13)).cuda()
    res = {}
    for hook in (3, mode):
        old = lib.tmb_tv_set_simple_kernels(hook)
        try:
            res[hook] = [PD_TV_cupy(v, 5e-4, its, m, nn, 12.0, 0, False).cpu().numpy()
                         for its, nn, m in ((2, 1, 0), (7, 0, 0), (4, 1, 1))]
        finally:
            lib.tmb_tv_set_simple_kernels(old)
    for a, b in zip(res[mode], res[3]):
        assert np.isfinite(a).all() and rel_max(a, b) < 2e-6



@pytest.mark.parametrize("shape", [(9, 21, 244), (66, 37, 364), (40, 130, 8)])
@pytest.mark.parametrize("half", [False, True])
def test_rof_fast_normalisation_stays_on_the_exact_path(shape, half):
    """k_rof_tv3d_w's default arithmetic (nom * MUFU.RSQ: raw approximation, <= 2^-22.9 relative) against the round-1
    path behind hook 3 (correctly rounded square root + IEEE division, the reference's own sequence): after 30
    iterations at the benchmark's parameters the two agree to 5e-7 of the volume's range."""
    from tomobar_b200._lib import lib
    from tomobar_b200.regularisersCuPy import ROF_TV_cupy

    g = torch.Generator(device="cuda").manual_seed(sum(shape))
    v = torch.randn(shape, device="cuda", generator=g) * 0.02
    fast = ROF_TV_cupy(v, 3e-4, 30, 1e-3, 0, half)
    old = lib.tmb_tv_set_simple_kernels(3)
    try:
        exact = ROF_TV_cupy(v, 3e-4, 30, 1e-3, 0, half)
    finally:
        lib.tmb_tv_set_simple_kernels(old)
    assert torch.isfinite(fast).all()
    assert (fast - exact).abs().max().item() <= 5e-7 * exact.abs().max().item()
