"""The 3-D host-array goldens of the reference's RecToolsDIR (tests/test_RecToolsDIR.py:221-323) through
``tomobar_b200.methodsDIR.RecToolsDIR`` (numpy in, numpy out over the CUDA path)."""

import os

import numpy as np
import pytest
from numpy.testing import assert_allclose

pytestmark = [pytest.mark.gpu]

LABELS = ["angles", "detY", "detX"]


def _rec(data, angles, pad=0):
    from tomobar_b200.methodsDIR import RecToolsDIR

    return RecToolsDIR(DetectorsDimH=data.shape[2], DetectorsDimH_pad=pad, DetectorsDimV=data.shape[1],
                       CenterRotOffset=0.0, AnglesVec=angles, ObjSize=data.shape[2], device_projector="gpu")


@pytest.mark.parametrize("pad", [0, 20])
def test_backproj3d(scan, pad):  # :221-262
    data, angles = scan
    bp = _rec(data, angles, pad).BACKPROJ(data, data_axes_labels_order=LABELS)
    assert_allclose(bp.min(), -3.8901403, rtol=1e-6)
    assert_allclose(bp.max(), 350.38193, rtol=1e-6)
    assert bp.dtype == np.float32 and bp.shape == (128, 160, 160)


@pytest.mark.parametrize("pad,lo,hi", [(0, -0.014693323, 0.0340156), (20, -0.013320876, 0.03534868)])
def test_fbp3d(scan, pad, lo, hi):  # :265-302
    data, angles = scan
    rec = _rec(data, angles, pad).FBP(data, data_axes_labels_order=LABELS)
    assert_allclose(rec.min(), lo, rtol=3e-6)
    assert_allclose(rec.max(), hi, rtol=3e-6)
    assert rec.dtype == np.float32 and rec.shape == (128, 160, 160)


def test_forwproj_then_default_axes(scan):
    """FORWPROJ returns [detY, angles, detX] by default and BACKPROJ / FBP accept it without labels."""
    data, angles = scan
    R = _rec(data, angles)
    vol = np.ones((128, 160, 160), np.float32)
    sino = R.FORWPROJ(vol)
    assert sino.shape == (128, 180, 160) and sino.dtype == np.float32
    assert_allclose(sino.min(), 67.27458, rtol=2e-6)   # tests/test_RecToolsDIRCuPy.py:691-692
    assert_allclose(sino.max(), 225.27428, rtol=2e-6)
    a = R.FBP(np.ascontiguousarray(data.swapaxes(0, 1)))          # already [detY, angles, detX]
    b = R.FBP(data, data_axes_labels_order=LABELS)
    assert np.array_equal(a, b)


def test_errors(scan):
    from tomobar_b200.methodsDIR import RecToolsDIR

    data, angles = scan
    with pytest.raises(ValueError):
        RecToolsDIR(160, 0, 128, 0.0, angles, 160, device_projector="cpu")
    with pytest.raises(NotImplementedError):
        RecToolsDIR(160, 0, None, 0.0, angles, 160)
    with pytest.raises(ValueError):
        _rec(data, angles).FBP(data.astype(np.float64), data_axes_labels_order=LABELS)


def test_fourier_inv_estimate_against_the_allocator():
    """The dry-run estimate (DeviceMemStack protocol) against torch's own peak statistic."""
    import torch

    from tomobar_b200.methodsDIR_CuPy import RecToolsDIRCuPy
    from tomobar_b200.supp.memory_estimator_helpers import DeviceMemStack

    nz, nproj, n = 64, 600, 1024
    angles = np.linspace(0, np.pi, nproj, endpoint=False).astype(np.float32)
    R = RecToolsDIRCuPy(n, 0, nz, 0.0, angles, n, device_projector=0)
    with DeviceMemStack() as st:
        assert R.FOURIER_INV((nz, nproj, n), data_dtype=np.float32) == (nz, n, n)
    data = torch.rand((nz, nproj, n), device="cuda")
    R.FOURIER_INV(data)  # cuFFT plans and the like are created on the first call
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    before = torch.cuda.memory_allocated()
    R.FOURIER_INV(data)
    torch.cuda.synchronize()
    measured = torch.cuda.max_memory_allocated() - before + data.numel() * 4
    assert 0.6 * st.highwater <= measured <= 1.05 * st.highwater, (measured, st.highwater)
