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
    assert RecToolsDIR(160, 0, None, 0.0, angles, 160).geom == "2D"
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


# ---- the 2-D geometry (AstraTools2D semantics: y-up images) -- tests/test_RecToolsDIR.py:13-169 ----------------
def _rec2d(angles, pad=0):
    from tomobar_b200.methodsDIR import RecToolsDIR

    return RecToolsDIR(DetectorsDimH=160, DetectorsDimH_pad=pad, DetectorsDimV=None, CenterRotOffset=0.0,
                       AnglesVec=angles, ObjSize=160, device_projector="gpu")


def test_backproj2d(scan):  # :13-34 (value ranges are all the reference pins) + the y-up relation to the 3-D path
    data, angles = scan
    data2d = np.ascontiguousarray(data[:, 60, :])
    bp = _rec2d(angles).BACKPROJ(data2d, data_axes_labels_order=["angles", "detX"])
    assert 22 <= bp.min() <= 25 and 130 <= bp.max() <= 150
    assert bp.dtype == np.float32 and bp.shape == (160, 160)
    bp3 = _rec(data, angles).BACKPROJ(data, data_axes_labels_order=LABELS)[60]
    assert np.array_equal(bp, bp3[::-1])  # ASTRA's 2-D image is the vertical flip of a slice of the 3-D volume
    swapped = _rec2d(angles).BACKPROJ(np.ascontiguousarray(data2d.T), data_axes_labels_order=["detX", "angles"])
    assert np.array_equal(bp, swapped)


def test_forwproj2d(scan):  # :37-62
    _, angles = scan
    R = _rec2d(angles)
    phantom = np.ones((160, 160), np.float32)
    fp = R.FORWPROJ(phantom, data_axes_labels_order=["angles", "detX"])
    assert 60 <= fp.min() <= 75 and 200 <= fp.max() <= 300
    assert fp.dtype == np.float32 and fp.shape == (180, 160)
    assert R.FORWPROJ(phantom, data_axes_labels_order=["detX", "angles"]).shape == (160, 180)
    rng = np.random.default_rng(3)
    img = rng.random((160, 160)).astype(np.float32)
    vol = np.zeros((128, 160, 160), np.float32)
    vol[7] = img[::-1]
    from tomobar_b200.methodsDIR import RecToolsDIR

    fp3 = RecToolsDIR(160, 0, 128, 0.0, angles, 160).FORWPROJ(vol)[7]
    assert np.array_equal(R.FORWPROJ(img), fp3)


@pytest.mark.parametrize("pad", [0, 20])
def test_fbp2d(scan, pad):  # :65-110
    data, angles = scan
    rec = _rec2d(angles, pad).FBP(np.ascontiguousarray(data[:, 60, :]), data_axes_labels_order=["angles", "detX"])
    assert rec.min() <= -0.0001 and rec.max() >= 0.001
    assert rec.dtype == np.float32 and rec.shape == (160, 160)


@pytest.mark.parametrize("filters_type", ["shepp-logan", "cosine", "hamming"])
@pytest.mark.parametrize("filter_d", [None, 0.1, 1.0])
def test_fbp2d_filters(scan, filters_type, filter_d):  # :112-141
    data, angles = scan
    rec = _rec2d(angles).FBP(np.ascontiguousarray(data[:, 60, :]), data_axes_labels_order=["angles", "detX"],
                             filter_type=filters_type, filter_parameter=None, filter_d=filter_d)
    assert rec.min() <= -0.0001 and rec.max() >= 0.001
    assert rec.dtype == np.float32 and rec.shape == (160, 160)


def test_fbp2d_mask_and_agreement_with_the_sinc_fbp(scan):  # :144-169
    data, angles = scan
    data2d = np.ascontiguousarray(data[:, 60, :])
    rec = _rec2d(angles).FBP(data2d, data_axes_labels_order=["angles", "detX"], recon_mask_radius=0.85)
    assert rec.min() <= -0.0001 and rec.max() >= 0.001
    assert np.sum(rec == 0) == 11963
    # the ram-lak FBP and the 3-D path's sinc-filter FBP (a = 1.1) are two filters for the same inversion
    full = _rec2d(angles).FBP(data2d)
    sinc = _rec(data, angles).FBP(data, data_axes_labels_order=LABELS)[60][::-1]
    assert np.corrcoef(full.ravel(), sinc.ravel())[0, 1] > 0.97
    assert 0.7 < np.abs(full).sum() / np.abs(sinc).sum() < 1.4


def test_fourier_inv_estimate_at_config4_within_5_percent():
    """BASELINE.json config 4 (2048^2 x 128, 2000 angles): the dry-run estimate against torch's measured peak
    (tools/check_estimator.py)."""
    import torch

    from tomobar_b200.methodsDIR_CuPy import RecToolsDIRCuPy
    from tomobar_b200.supp.memory_estimator_helpers import DeviceMemStack

    nz, nproj, n = 128, 2000, 2048
    if torch.cuda.get_device_properties(0).total_memory < 40e9:
        pytest.skip("needs ~17 GB of device memory")
    angles = np.linspace(0, np.pi, nproj, endpoint=False).astype(np.float32)
    R = RecToolsDIRCuPy(n, 0, nz, 0.0, angles, n, device_projector=0)
    with DeviceMemStack() as st:
        assert R.FOURIER_INV((nz, nproj, n), data_dtype=np.float32) == (nz, n, n)
    data = torch.rand((nz, nproj, n), device="cuda")
    R.FOURIER_INV(data)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
    before = torch.cuda.memory_allocated()
    R.FOURIER_INV(data)
    torch.cuda.synchronize()
    measured = torch.cuda.max_memory_allocated() - before + data.numel() * 4
    assert abs(measured / st.highwater - 1.0) < 0.05, (measured, st.highwater)
    del data
    torch.cuda.empty_cache()
