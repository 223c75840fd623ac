"""Loop-level parity of every reconstruction method of RecToolsIRCuPy with the oracle's restatement
of the reference loops (methodsIR_CuPy.py:128-667) on a small synthetic problem: Landweber, SIRT,
CGLS, OSEM / MLEM (+ TV), FISTA with the KL data term, ADMM with PWLS."""

import numpy as np
import pytest
import torch

from conftest import rel_max

pytestmark = pytest.mark.gpu

NZ, N, NA = 6, 48, 60


@pytest.fixture(scope="module")
def problem(oracle):
    rng = np.random.default_rng(3)
    angles = np.linspace(0, np.pi, NA, endpoint=False).astype(np.float32)
    yy, xx = np.mgrid[:N, :N]
    disc = (((xx - N / 2 + 3) ** 2 + (yy - N / 2 - 2) ** 2) < (0.3 * N) ** 2).astype(np.float32)
    vol = np.stack([disc * (1 + 0.15 * z) for z in range(NZ)]).astype(np.float32) * 0.04
    b = oracle.RecIR(N, 0, NZ, 0.0, angles, N, None)._Ax(vol)
    b = np.maximum(b + 0.01 * rng.standard_normal(b.shape).astype(np.float32), 1e-3).astype(np.float32)
    return angles, b


def _pair(oracle, angles, os_n):
    from tomobar_b200.methodsIR_CuPy import RecToolsIRCuPy

    return oracle.RecIR(N, 0, NZ, 0.0, angles, N, os_n), RecToolsIRCuPy(N, 0, NZ, 0.0, angles, N, 0, os_n)


def _data(b, **kw):
    d = {"projection_data": torch.from_numpy(b).cuda()}
    d.update(kw)
    return d


def test_landweber_sirt_cgls(oracle, problem):
    angles, b = problem
    ref, rec = _pair(oracle, angles, None)
    got = rec.Landweber(_data(b), {"iterations": 12, "tau_step_lanweber": 2e-4, "recon_mask_radius": None})
    assert rel_max(got.cpu().numpy(), ref.Landweber(b, 12, 2e-4, mask_radius=None)) < 2e-5
    got = rec.SIRT(_data(b), {"iterations": 6, "nonnegativity": True})
    assert rel_max(got.cpu().numpy(), ref.SIRT(b, 6, nonneg=True)) < 2e-5
    got = rec.CGLS(_data(b), {"iterations": 6, "recon_mask_radius": None})
    assert rel_max(got.cpu().numpy(), ref.CGLS(b, 6, mask_radius=None)) < 2e-4  # inner products: fp32 order


@pytest.mark.parametrize("os_n", [None, 5])
@pytest.mark.parametrize("tv", [False, True])
def test_osem(oracle, problem, os_n, tv):
    angles, b = problem
    ref, rec = _pair(oracle, angles, os_n)
    reg = {"method": "PD_TV", "regul_param": 2e-4, "iterations": 5} if tv else None
    want = ref.OSEM(b, 4, regularisation=dict(reg) if reg else None, mask_radius=None)
    got = rec.OSEM(_data(b), {"iterations": 4, "recon_mask_radius": None}, dict(reg) if reg else None)
    assert rel_max(got.cpu().numpy(), want) < 5e-5


@pytest.mark.parametrize("os_n", [None, 4])
def test_fista_kl_and_admm_pwls(oracle, problem, os_n):
    angles, b = problem
    ref, rec = _pair(oracle, angles, os_n)
    L = 3000.0 if os_n is None else 800.0
    rof = {"method": "ROF_TV", "regul_param": 3e-4, "iterations": 6, "time_marching_step": 1e-3}
    want = ref.FISTA(b, 4, lipschitz_const=L, regularisation=dict(rof), nonneg=True, fidelity="KL", mask_radius=None)
    got = rec.FISTA(_data(b, data_fidelity="KL"),
                    {"iterations": 4, "lipschitz_const": L, "nonnegativity": True, "recon_mask_radius": None}, dict(rof))
    assert rel_max(got.cpu().numpy(), want) < 5e-5
    pd = {"method": "PD_TV", "regul_param": 3e-4, "iterations": 6, "methodTV": 1}
    want = ref.ADMM(b, 4, lipschitz_const=L, regularisation=dict(pd), fidelity="PWLS", rho=1.0, relax=1.6,
                    mask_radius=None)
    got = rec.ADMM(_data(b, data_fidelity="PWLS"),
                   {"iterations": 4, "lipschitz_const": L, "ADMM_rho_const": 1.0, "ADMM_relax_par": 1.6,
                    "recon_mask_radius": None}, dict(pd))
    assert rel_max(got.cpu().numpy(), want) < 5e-5
