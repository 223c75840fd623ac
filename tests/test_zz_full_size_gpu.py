"""Size-independent properties at BASELINE.json's FULL configuration sizes (the oracle cannot follow there):
linearity and slice independence of the projector pair, fixed points and shift equivariance of the TV prox,
agreement of the kernel families, z-block invariance of FOURIER_INV, robust data terms reducing to LS.
Runs last (file name) because it needs tens of GB and a few seconds; every case frees its memory."""

import numpy as np
import pytest
import torch

from conftest import rel_max

pytestmark = pytest.mark.gpu


def _angles(na):
    return np.linspace(0, np.pi, na, endpoint=False).astype(np.float32)


def _free():
    torch.cuda.synchronize()
    torch.cuda.empty_cache()


def _need_gb(gb):
    free, _ = torch.cuda.mem_get_info()
    if free < gb * 1e9:
        pytest.skip(f"needs {gb} GB of free device memory")


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def test_config2_projector_pair_1024x1024x256_900_angles():
    """BASELINE config 2: OS = 6 subsets of 150 angles."""
    from tomobar_b200.projector import ProjTools3D

    _need_gb(12)
    nz, n, na = 256, 1024, 900
    P = ProjTools3D(n, 0, nz, _angles(na), 0.0, n, "gpu", 0, 6)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(nz, n, n, device="cuda", generator=g)
    y = torch.randn(nz, n, n, device="cuda", generator=g)
    fx, fy = P._forwprojOSCuPy(x, 2), P._forwprojOSCuPy(y, 2)
    assert fx.shape == (nz, 150, n)
    assert _rel(P._forwprojOSCuPy(2.0 * x - 3.0 * y, 2), 2.0 * fx - 3.0 * fy) < 1e-4        # linearity
    assert torch.count_nonzero(P._forwprojOSCuPy(torch.zeros_like(x), 2)) == 0
    assert torch.equal(P._forwprojOSCuPy(x.flip(0).contiguous(), 2), fx.flip(0))              # slices are independent
    bx = P._backprojOSCuPy(fx, 2)
    assert bx.shape == (nz, n, n) and torch.isfinite(bx).all()
    assert _rel(P._backprojOSCuPy(2.0 * fx - 3.0 * fy, 2), 2.0 * bx - 3.0 * P._backprojOSCuPy(fy, 2)) < 1e-4
    assert torch.equal(P._backprojOSCuPy(fx.flip(0).contiguous(), 2), bx.flip(0))
    del P, x, y, fx, fy, bx
    _free()


def test_config2_tv_prox_256x1024x1024():
    """PD_TV x 50 inner iterations (config 2's prox) and ROF_TV: constants are fixed points, the prox
    commutes with adding a constant, the fused and the one-iteration-per-launch kernels agree."""
    from tomobar_b200._lib import lib
    from tomobar_b200.regularisersCuPy import PD_TV_cupy, ROF_TV_cupy

    _need_gb(20)
    shape = (256, 1024, 1024)
    const = torch.full(shape, 0.37, device="cuda")
    assert _rel(PD_TV_cupy(const, 3e-4, 50, 0, 0, 12.0, 0, False), const) < 2e-6
    assert _rel(ROF_TV_cupy(const, 3e-4, 10, 1e-3, 0, False), const) < 2e-6
    del const
    g = torch.Generator(device="cuda").manual_seed(1)
    v = torch.rand(shape, device="cuda", generator=g) * 0.02
    a = PD_TV_cupy(v, 3e-4, 50, 0, 0, 12.0, 0, False)
    assert torch.isfinite(a).all() and a.std() < v.std()                          # it does smooth
    b = PD_TV_cupy(v + 0.25, 3e-4, 50, 0, 0, 12.0, 0, False)
    assert ((b - 0.25) - a).abs().max().item() < 2e-5                             # shift equivariance (values ~0.26)
    del b
    old = lib.tmb_tv_set_simple_kernels(3)
    try:
        c = PD_TV_cupy(v, 3e-4, 50, 0, 0, 12.0, 0, False)
    finally:
        lib.tmb_tv_set_simple_kernels(old)
    assert _rel(a, c) < 2e-6                                                       # fused pairs vs single iterations
    del a, c, v
    _free()


def test_headline_tv_prox_512x2048x2048():
    """The bench volume: fused pairs of iterations against single iterations, 5 iterations (odd tail)."""
    from tomobar_b200._lib import lib
    from tomobar_b200.regularisersCuPy import PD_TV_cupy

    _need_gb(120)
    g = torch.Generator(device="cuda").manual_seed(2)
    v = torch.rand((512, 2048, 2048), device="cuda", generator=g) * 0.02
    a = PD_TV_cupy(v, 3e-4, 5, 0, 1, 12.0, 0, False)
    old = lib.tmb_tv_set_simple_kernels(3)
    try:
        c = PD_TV_cupy(v, 3e-4, 5, 0, 1, 12.0, 0, False)
    finally:
        lib.tmb_tv_set_simple_kernels(old)
    assert torch.isfinite(a).all() and _rel(a, c) < 2e-6
    del a, c, v
    _free()


def test_config4_fourier_inv_2048x2048x128_2000_angles():
    """BASELINE config 4: a z-block of the projections reconstructs to the same slices as the whole stack
    (the reference's own property test, tests/test_RecToolsDIRCuPy.py:253-288, at full size)."""
    from tomobar_b200.methodsDIR_CuPy import RecToolsDIRCuPy

    _need_gb(40)
    nz, na, n = 128, 2000, 2048
    g = torch.Generator(device="cuda").manual_seed(3)
    data = torch.rand((nz, na, n), device="cuda", generator=g)                     # [detY, angles, detX]
    whole = RecToolsDIRCuPy(n, 0, nz, 0.0, _angles(na), n, device_projector=0).FOURIER_INV(data)
    assert whole.shape == (nz, n, n) and torch.isfinite(whole).all()
    block = RecToolsDIRCuPy(n, 0, 32, 0.0, _angles(na), n, device_projector=0).FOURIER_INV(data[64:96].contiguous())
    assert _rel(block, whole[64:96]) < 5e-5
    del data, whole, block
    _free()


def test_config5_robust_terms_reduce_to_ls_96x1536x1536_1500_angles():
    """BASELINE config 5, one GPU's shard (384 / 4 slices): a Huber threshold no residual reaches and a ring
    model switched off leave the LS iteration unchanged."""
    from tomobar_b200.methodsIR_CuPy import RecToolsIRCuPy

    _need_gb(20)
    nz, n, na = 96, 1536, 1500
    g = torch.Generator(device="cuda").manual_seed(4)
    b = torch.rand((nz, na, n), device="cuda", generator=g)
    alg = {"iterations": 1, "lipschitz_const": 4.0e5, "nonnegativity": True, "recon_mask_radius": None}
    rec = RecToolsIRCuPy(n, 0, nz, 0.0, _angles(na), n, 0, None)
    ls = rec.FISTA({"projection_data": b}, dict(alg))
    hub = rec.FISTA({"projection_data": b, "huber_threshold": 1.0e30}, dict(alg))
    assert torch.isfinite(ls).all() and ls.abs().max() > 0
    assert _rel(hub, ls) < 1e-6
    del b, ls, hub, rec
    _free()
