"""FOURIER_INV: the reference's pinned goldens, z-block invariance, odd/even shape matrix
(reference tests/test_RecToolsDIRCuPy.py:225-468) and kernel-level parity with the reference's
own fft_us_kernels.cu kernels run from oracle/_ref."""

import math

import numpy as np
import pytest
import torch
from numpy.testing import assert_allclose

import ref_kernels as R
from conftest import rel_l2, rel_max

pytestmark = pytest.mark.gpu
LABELS = ["angles", "detY", "detX"]


def _dir(angles, detX, detY, obj=None, pad=0):
    from tomobar_b200.methodsDIR_CuPy import RecToolsDIRCuPy

    return RecToolsDIRCuPy(DetectorsDimH=detX, DetectorsDimH_pad=pad, DetectorsDimV=detY, CenterRotOffset=0.0,
                           AnglesVec=angles, ObjSize=detX if obj is None else obj, device_projector=0)


def test_fourier_inv_golden(scan):  # tests/test_RecToolsDIRCuPy.py:225-250
    data, angles = scan
    rec = _dir(angles, 160, 128).FOURIER_INV(torch.from_numpy(data).cuda(), data_axes_labels_order=LABELS,
                                             recon_mask_radius=2.0).cpu().numpy()
    assert_allclose(rec.min(), -0.0372409, atol=1e-5)
    assert_allclose(rec.max(), 0.1035610, atol=1e-4)
    assert rec.dtype == np.float32 and rec.shape == (128, 160, 160)


@pytest.mark.parametrize("blocks", [2, 8, 32])
def test_fourier_inv_vert_blocks(scan, blocks):  # :253-288
    data, angles = scan
    d = torch.from_numpy(data).cuda()
    R_ = _dir(angles, 160, blocks)
    out = torch.empty((128, 160, 160), device="cuda")
    for s in range(0, 128, blocks):
        out[s:s + blocks] = R_.FOURIER_INV(d[:, s:s + blocks, :], recon_mask_radius=2.0, data_axes_labels_order=LABELS)
    rec = out.cpu().numpy()
    assert_allclose(rec.min(), -0.0372409, atol=1e-5)
    assert_allclose(rec.max(), 0.1035610, atol=1e-4)


def test_fourier_inv_last_block_odd(scan):  # :291-337
    data, angles = scan
    d = torch.from_numpy(data).cuda()
    a = _dir(angles, 160, 125).FOURIER_INV(d[:, :125, :], recon_mask_radius=2.0, data_axes_labels_order=LABELS)
    b = _dir(angles, 160, 3).FOURIER_INV(d[:, 125:, :], recon_mask_radius=2.0, data_axes_labels_order=LABELS)
    rec = torch.cat((a, b)).cpu().numpy()
    assert rec.shape == (128, 160, 160)
    assert_allclose(rec.min(), -0.0372409, atol=1e-5)
    assert_allclose(rec.max(), 0.1035610, atol=1e-4)


@pytest.mark.parametrize("detX,obj", [(341, 340), (342, 341), (342, 342), (341, 341)])
def test_fourier_inv_odd_even_shapes(detX, obj):  # :340-440 (sizes reduced 4x)
    rng = np.random.default_rng(0)
    data = rng.integers(7515, 37624, size=(225, 3, detX)).astype(np.float32)
    angles = np.linspace(0, math.pi, data.shape[0])
    rec = _dir(angles, detX, 3, obj).FOURIER_INV(torch.from_numpy(data).cuda(), data_axes_labels_order=LABELS)
    assert rec.dtype == torch.float32 and tuple(rec.shape) == (3, obj, obj)
    assert torch.isfinite(rec).all()


@pytest.mark.skipif(not R.available(), reason="oracle/_ref cubins not built")
@pytest.mark.parametrize("case", [(16, 180, 160, 160, 0), (6, 120, 97, 96, 0), (5, 90, 130, 100, 3), (4, 400, 256, 256, 0)])
def test_fourier_inv_vs_reference_kernels(scan, case):
    nz, na, detX, obj, pad = case
    data, angles_scan = scan
    if (na, detX) == (180, 160):
        d = torch.from_numpy(np.ascontiguousarray(np.swapaxes(data[:, 40:40 + nz, :], 0, 1))).cuda()
        angles = angles_scan
    else:
        g = torch.Generator(device="cuda").manual_seed(na)
        d = torch.rand((nz, na, detX), device="cuda", generator=g)
        angles = np.linspace(0, math.pi, na, endpoint=False).astype(np.float32)
        if na == 400:
            angles = np.linspace(0, 2 * math.pi, na, endpoint=False).astype(np.float32)  # 360-degree scan
    ref = R.ref_FOURIER_INV(d, angles, obj, 0.0, pad).cpu().numpy()
    from tomobar_b200.methodsDIR_CuPy import RecToolsDIRCuPy

    got = RecToolsDIRCuPy(detX, pad, nz, 0.0, angles, obj, device_projector=0).FOURIER_INV(d).cpu().numpy()
    assert got.shape == ref.shape
    assert rel_l2(got, ref) < 1e-5, rel_l2(got, ref)
    assert rel_max(got, ref) < 1e-4, rel_max(got, ref)


@pytest.mark.skipif(not R.available(), reason="oracle/_ref cubins not built")
@pytest.mark.parametrize("center_size", [0, 100, 192, 256, 300])
def test_fourier_inv_scatter_branches_vs_reference_kernels(scan, center_size):
    """The non-default branches (methodsDIR_CuPy.py:759-835): center_size < 192 scatters every polar sample onto the
    grid (gather_kernel), 192 <= center_size < 2n gathers a centre square and scatters the rest
    (gather_kernel_partial + gather_kernel_center), against a chain of the reference's own kernels."""
    data, angles = scan
    nz, detX = 6, 160
    d = torch.from_numpy(np.ascontiguousarray(np.swapaxes(data[:, 50:50 + nz, :], 0, 1))).cuda()
    ref = R.ref_FOURIER_INV(d, angles, detX, center_size=center_size).cpu().numpy()
    from tomobar_b200.methodsDIR_CuPy import RecToolsDIRCuPy

    got = RecToolsDIRCuPy(detX, 0, nz, 0.0, angles, detX, device_projector=0).FOURIER_INV(d, center_size=center_size)
    got = got.cpu().numpy()
    assert got.shape == ref.shape
    assert rel_l2(got, ref) < 2e-5, rel_l2(got, ref)   # atomic adds: the order of the sums is not fixed
    assert rel_max(got, ref) < 2e-4, rel_max(got, ref)
    full = RecToolsDIRCuPy(detX, 0, nz, 0.0, angles, detX, device_projector=0).FOURIER_INV(d).cpu().numpy()
    # scatter and gather formulations are NOT numerically equal (3.7 % apart in rel-L2 there)
    assert rel_l2(got, full) < 0.1


@pytest.mark.parametrize("n,na,nz,span,center", [(362, 241, 10, math.pi, None), (256, 180, 6, 2 * math.pi, None),
                                                  (200, 97, 4, math.pi, 224)])
def test_gather_variants_are_bit_identical(n, na, nz, span, center):
    """k_fi_gather_w (the default: a warp walks the polar lines of its 8 x 4 patch of grid points in lock step) and
    k_fi_gather_s (hook 2: the samples of a 16 x 8 tile staged in shared memory, the tile's angle range walked in
    batches) visit every (point, line, sample) of k_fi_gather (hook 1: every thread walks its own lines) in the same
    order: the grids are bit-identical, on the whole grid and on a centre square."""
    from tomobar_b200._lib import lib, check
    from tomobar_b200._tensors import ptr

    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream(dev).cuda_stream
    nz2 = nz // 2
    theta = torch.as_tensor(-np.linspace(0, span, na, endpoint=False), dtype=torch.float32, device=dev)
    sorted_theta, sorted_idx = torch.sort(theta)
    sorted_idx = sorted_idx.to(torch.int32)
    g = torch.Generator(device="cuda").manual_seed(n)
    datac = torch.view_as_complex(torch.randn((nz2, na, n, 2), device=dev, generator=g))
    mu = -np.log(1e-4) / (2 * n * n)
    m = int(np.ceil(2 * n * 1 / np.pi * np.sqrt(-mu * np.log(1e-4) + (mu * n) * (mu * n) / 4)))
    out = {}
    for mode in (1, 2, 3, 0):
        fde = torch.zeros((nz2, 2 * n, 2 * n), dtype=torch.complex64, device=dev)
        old = lib.tmb_fi_set_gather(mode)
        try:
            if center is None:
                check(lib.tmb_fi_gather(ptr(datac), ptr(fde), ptr(theta), ptr(sorted_theta), ptr(sorted_idx), m,
                                        float(np.float32(mu)), n, na, nz2, st), "tmb_fi_gather")
            else:
                check(lib.tmb_fi_gather_center(ptr(datac), ptr(fde), ptr(theta), ptr(sorted_theta), ptr(sorted_idx), m,
                                               float(np.float32(mu)), n, na, nz2, center, st), "tmb_fi_gather_center")
        finally:
            lib.tmb_fi_set_gather(old)
        out[mode] = torch.view_as_real(fde)
    assert torch.isfinite(out[3]).all() and out[1].abs().max() > 0
    for mode in (2, 3, 0):
        assert torch.equal(out[1], out[mode])


@pytest.mark.parametrize("case", [(16, 180, 160, 160, 0, 0.0), (6, 120, 97, 96, 0, 2.5), (4, 90, 130, 100, 3, -1.25),
                                  (8, 64, 256, 256, 0, 0.0)])
@pytest.mark.parametrize("filter_type", ["shepp", "hann"])
@pytest.mark.parametrize("pow2", [True, False])
def test_fourier_filter_slice_pairs_match_per_slice_filter(case, filter_type, pow2):
    """STEP 0 of FOURIER_INV (methodsDIR_CuPy.py:449-545) on complex slice-pair rows (one c2c transform pair per slice
    pair, two-sided filter with the real Nyquist bin irfft implies) against the rfft / irfft per slice: the filtered,
    packed projections agree to fp32 FFT rounding, with a shifted rotation axis and without power-of-two oversampling."""
    from tomobar_b200.methodsDIR_CuPy import RecToolsDIRCuPy

    nz, na, detX, obj, pad, cor = case
    g = torch.Generator(device="cuda").manual_seed(na + detX)
    d = torch.rand((nz, na, detX + detX % 2), device="cuda", generator=g)
    angles = np.linspace(0, math.pi, na, endpoint=False).astype(np.float32)
    T = RecToolsDIRCuPy(detX, pad, nz, cor, angles, obj, device_projector=0)
    n = d.shape[-1] + 2 * pad
    out = []
    for pairs in (True, False):
        T._FILTER_SLICE_PAIRS = pairs
        datac = torch.empty((nz // 2, na, n), dtype=torch.complex64, device="cuda")
        assert T._fourier_filter(d, d.shape[-1], n, pow2, 4, filter_type, 1.0, pack_into=datac) is None
        out.append(torch.view_as_real(datac).cpu().numpy())
    assert rel_l2(out[0], out[1]) < 2e-6, rel_l2(out[0], out[1])
    assert rel_max(out[0], out[1]) < 2e-5, rel_max(out[0], out[1])


@pytest.mark.parametrize("pairs", [True, False])
def test_fourier_filter_8192_point_rows_vs_float64(pairs):
    """STEP 0 at config 4's row length (2048 detector pixels oversampled to 8192) against a float64 rfft / irfft with
    numpy's irfft semantics (imaginary part of the Nyquist bin ignored).  The phase ramp of a centred rotation axis makes
    that bin purely imaginary; cuFFT's c2r used it at this size (7.7e-3 relative) until the bin was made real."""
    from tomobar_b200.fourier import calc_filter
    from tomobar_b200.methodsDIR_CuPy import RecToolsDIRCuPy

    n, nz, na = 2048, 4, 24
    angles = np.linspace(0, math.pi, na, endpoint=False).astype(np.float32)
    T = RecToolsDIRCuPy(n, 0, nz, 0.0, angles, n, device_projector=0)
    T._FILTER_SLICE_PAIRS = pairs
    g = torch.Generator(device="cuda").manual_seed(5)
    d = torch.rand((nz, na, n), device="cuda", generator=g)
    datac = torch.empty((nz // 2, na, n), dtype=torch.complex64, device="cuda")
    T._fourier_filter(d, n, n, True, 4, "shepp", 1.0, pack_into=datac)
    over = 8192
    pm = over // 2 - n // 2
    w64 = torch.as_tensor(calc_filter(over, "shepp", 1.0), device="cuda").double() * torch.exp(
        (-2 * np.pi * 1j * 0.5) * torch.fft.rfftfreq(over, device="cuda").double())
    torch.view_as_real(w64)[over // 2, 1] = 0.0
    x = torch.nn.functional.pad(d.double(), (pm, over - pm - n), mode="replicate")
    y = torch.fft.irfft(w64 * torch.fft.rfft(x, dim=2), n=over, dim=2)[:, :, pm:pm + n]
    sgn = torch.where(torch.arange(n, device="cuda") % 2 == 1, 1.0, -1.0)
    ref = torch.complex(y[0::2] * sgn, y[1::2] * sgn)
    err = float((datac - ref).norm() / ref.norm())
    assert err < 2e-6, err


@pytest.mark.parametrize("n,na,nz2", [(128, 90, 8), (96, 64, 16), (64, 50, 24), (80, 50, 40), (200, 97, 32)])
def test_slice_pair_gather_is_bit_identical(n, na, nz2):
    """tmb_fi_scale_sign_pairs -> tmb_fi_gather_pairs (polar samples stored as slice pairs, one 128-bit load per two slices,
    8 or 16 complex slices per thread, no per-slice predicates) against tmb_fi_scale_sign -> k_fi_gather (hook 1: planar
    samples, every thread walks its own lines, predicated scalar loads): same (point, line, sample) visits in the same
    order, the same grid bit for bit -- as are the whole-chunk (FULL) planar variants at 4 / 8 / 16 slices per thread
    (gather_kernel_center, fft_us_kernels.cu:468-527; c1dfftshift :559-586)."""
    from tomobar_b200._lib import lib, check
    from tomobar_b200._tensors import ptr

    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream(dev).cuda_stream
    theta = torch.as_tensor(-np.linspace(0, math.pi, na, endpoint=False), dtype=torch.float32, device=dev)
    sorted_theta, sorted_idx = torch.sort(theta)
    sorted_idx = sorted_idx.to(torch.int32)
    g = torch.Generator(device="cuda").manual_seed(n + nz2)
    datac = torch.view_as_complex(torch.randn((nz2, na, n, 2), device=dev, generator=g))
    dataz = torch.empty_like(datac)
    c = float(np.float32(4 / n))
    check(lib.tmb_fi_scale_sign_pairs(ptr(datac), ptr(dataz), c, n, na, nz2, st), "tmb_fi_scale_sign_pairs")
    check(lib.tmb_fi_scale_sign(ptr(datac), c, n, na, nz2, st), "tmb_fi_scale_sign")
    z = torch.view_as_real(dataz).view(nz2 // 2, na, n, 2, 2)
    assert torch.equal(z[:, :, :, 0], torch.view_as_real(datac[0::2])) and torch.equal(z[:, :, :, 1], torch.view_as_real(datac[1::2]))
    mu = -np.log(1e-4) / (2 * n * n)
    m = int(np.ceil(2 * n * 1 / np.pi * np.sqrt(-mu * np.log(1e-4) + (mu * n) * (mu * n) / 4)))

    def gather(fn, src, mode=0, sc=0):
        fde = torch.full((nz2, 2 * n, 2 * n), float("nan"), dtype=torch.complex64, device=dev)
        old_m, old_s = lib.tmb_fi_set_gather(mode), lib.tmb_fi_set_slices_per_thread(sc)
        try:
            check(fn(ptr(src), ptr(fde), ptr(theta), ptr(sorted_theta), ptr(sorted_idx), m, float(np.float32(mu)), n, na,
                     nz2, st), "gather")
        finally:
            lib.tmb_fi_set_gather(old_m), lib.tmb_fi_set_slices_per_thread(old_s)
        return torch.view_as_real(fde)

    ref = gather(lib.tmb_fi_gather, datac, mode=1)
    assert torch.isfinite(ref).all() and ref.abs().max() > 0
    for sc in (0, 8, 16):
        assert torch.equal(ref, gather(lib.tmb_fi_gather_pairs, dataz, sc=sc)), sc
    for sc in (0, 4, 8, 16, 108):
        assert torch.equal(ref, gather(lib.tmb_fi_gather, datac, mode=3, sc=sc)), sc


@pytest.mark.parametrize("nz,na,detX", [(16, 90, 128), (32, 64, 96), (48, 50, 96), (80, 50, 100), (16, 120, 97)])
def test_fourier_inv_slice_pair_gather_matches_planar(nz, na, detX):
    """FOURIER_INV through the slice-pair layout against the planar one: bit-identical reconstructions.  (Detectors of
    96 pixels and more: below a 192-point grid the method scatters with atomic adds like the reference, methodsDIR_CuPy.py
    :761-779, the slice-pair gather is not involved and two calls differ in the last bit.)"""
    from tomobar_b200.methodsDIR_CuPy import RecToolsDIRCuPy

    g = torch.Generator(device="cuda").manual_seed(nz + na)
    d = torch.rand((nz, na, detX), device="cuda", generator=g)
    angles = np.linspace(0, math.pi, na, endpoint=False).astype(np.float32)
    T = RecToolsDIRCuPy(detX, 0, nz, 0.0, angles, detX, device_projector=0)
    out = []
    for pairs in (True, False):
        T._GATHER_SLICE_PAIRS = pairs
        out.append(T.FOURIER_INV(d))
    assert torch.isfinite(out[0]).all()
    assert torch.equal(out[0], out[1])
