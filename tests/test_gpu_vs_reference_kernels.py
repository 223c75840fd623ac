"""libtmb's TV / filter kernels against the reference's OWN kernels (oracle/_ref cubins built by
oracle/build_ref.sh from /root/reference) executed on the same GPU."""

import numpy as np
import pytest
import torch

import ref_kernels as R
from conftest import rel_max

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not R.available(), reason="oracle/_ref cubins not built")]


def _vol(shape, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    v = torch.randn(shape, device="cuda", generator=g) * 0.01
    v += (torch.rand(shape, device="cuda", generator=g) > 0.5).float() * 0.02
    return v


@pytest.mark.parametrize("shape", [(16, 70, 130), (80, 37, 200), (64, 129), (1, 50, 300), (70, 21, 132), (96, 18, 64),
                                   (9, 35, 388)])
@pytest.mark.parametrize("methodTV,nonneg", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("half", [False, True])
def test_pd_tv_vs_reference_kernel(shape, methodTV, nonneg, half):
    from tomobar_b200.regularisersCuPy import PD_TV_cupy

    v = _vol(shape, 3)
    ref = R.ref_PD_TV(v, 4e-4, 25, methodTV, nonneg, 12.0, half)
    out = PD_TV_cupy(v, 4e-4, 25, methodTV, nonneg, 12.0, 0, half)
    torch.cuda.synchronize()
    assert out.shape == ref.shape
    err = rel_max(out.cpu().numpy(), ref.cpu().numpy())
    # same arithmetic, possibly different FMA contraction; fp16 storage amplifies last-bit flips
    assert err < (2e-3 if half else 2e-6), err


@pytest.mark.parametrize("shape", [(16, 70, 130), (80, 37, 200), (64, 129)])
@pytest.mark.parametrize("half", [False, True])
def test_rof_tv_vs_reference_kernel(shape, half):
    from tomobar_b200.regularisersCuPy import ROF_TV_cupy

    v = _vol(shape, 4)
    ref = R.ref_ROF_TV(v, 3e-4, 30, 1e-3, half)
    out = ROF_TV_cupy(v, 3e-4, 30, 1e-3, 0, half)
    torch.cuda.synchronize()
    err = rel_max(out.cpu().numpy(), ref.cpu().numpy())
    assert err < (2e-3 if half else 2e-6), err


@pytest.mark.parametrize("n", [160, 200, 2048, 2560])
@pytest.mark.parametrize("cutoff", [0.35, 1.1])
def test_sinc_filter_vs_reference_kernel(n, cutoff):
    from tomobar_b200.fourier import sinc_filter

    ref = R.ref_filtersinc(n, cutoff, 1.0 / 180 / n)
    out = sinc_filter(n, cutoff, 1.0 / 180 / n, "cuda")
    torch.cuda.synchronize()
    assert rel_max(out.cpu().numpy(), ref.cpu().numpy()) < 1e-5
