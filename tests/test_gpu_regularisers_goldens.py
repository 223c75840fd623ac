"""The reference's tests/test_regularisers.py (pinned means of the TV-denoised test scan) through
the CUDA kernels, plus its unit-axis squeezing checks."""

import numpy as np
import pytest
import torch
from numpy.testing import assert_allclose

from tomobar_b200.regularisersCuPy import _squeeze_unit_axis


@pytest.mark.parametrize("shape,out_shape,flag,axis", [((100, 100), (100, 100), True, 0),
                                                       ((10, 100, 100), (10, 100, 100), False, 0),
                                                       ((1, 100, 100), (100, 100), True, 0),
                                                       ((16, 1, 100), (16, 100), True, 1)])
def test_check_if_input_2d_or_3d(shape, out_shape, flag, axis):  # tests/test_regularisers.py:9-39
    data, is2d, ind_axis = _squeeze_unit_axis(torch.zeros(shape))
    assert tuple(data.shape) == out_shape and is2d == flag and ind_axis == axis
    with pytest.raises(ValueError):
        _squeeze_unit_axis(torch.zeros((2, 2, 2, 2)))


@pytest.mark.gpu
def test_pd_tv_3d_mean(scan):  # :42-56
    from tomobar_b200.regularisersCuPy import PD_TV_cupy

    data = torch.from_numpy(scan[0]).cuda()
    den = PD_TV_cupy(data, regularisation_parameter=0.05, iterations=200, methodTV=0, nonneg=0, lipschitz_const=8,
                     gpu_id=0, half_precision=False)
    assert_allclose(float(den.mean()), 0.289258, atol=1e-4)
    assert den.shape == (180, 128, 160) and den.dtype == torch.float32


@pytest.mark.gpu
def test_pd_tv_2d_half_mean(scan):  # :59-73
    from tomobar_b200.regularisersCuPy import PD_TV_cupy

    data = torch.from_numpy(np.ascontiguousarray(scan[0][:, 64, :])).cuda()
    den = PD_TV_cupy(data, regularisation_parameter=0.05, iterations=200, methodTV=0, nonneg=0, lipschitz_const=8,
                     gpu_id=0, half_precision=True)
    assert_allclose(float(den.mean()), 0.2911671996116638, atol=1e-4)
    assert den.shape == (1, 180, 160) and den.dtype == torch.float32


@pytest.mark.gpu
def test_rof_tv_3d_mean(scan):  # :76-87
    from tomobar_b200.regularisersCuPy import ROF_TV_cupy

    data = torch.from_numpy(scan[0]).cuda()
    den = ROF_TV_cupy(data, regularisation_parameter=0.05, iterations=200, gpu_id=0, half_precision=False)
    assert_allclose(float(den.mean()), 0.289244, atol=1e-4)
    assert den.shape == (180, 128, 160) and den.dtype == torch.float32


@pytest.mark.gpu
def test_rof_tv_2d_mean(scan):  # :90-98
    from tomobar_b200.regularisersCuPy import ROF_TV_cupy

    data = torch.from_numpy(np.ascontiguousarray(scan[0][:, 64, :])).cuda()
    den = ROF_TV_cupy(data, regularisation_parameter=0.5, iterations=200, gpu_id=0, half_precision=False)
    assert_allclose(float(den.mean()), 0.29102084040641785, atol=1e-4)
    assert den.shape == (1, 180, 160) and den.dtype == torch.float32
