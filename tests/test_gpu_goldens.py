"""The reference's own pinned goldens, reproduced by the CUDA path through the public classes
on the reference's own scan.  Tolerances are the reference tests' where the restated ASTRA
model reaches them, otherwise the measured oracle-vs-golden gap."""

import numpy as np
import pytest
import torch
from numpy.testing import assert_allclose

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu_scan(scan):
    data, angles = scan
    return torch.from_numpy(data).cuda(), angles


def _ir(angles, detX, detY, pad=0, os_n=None):
    from tomobar_b200.methodsIR_CuPy import RecToolsIRCuPy

    return RecToolsIRCuPy(DetectorsDimH=detX, DetectorsDimH_pad=pad, DetectorsDimV=detY, CenterRotOffset=0.0,
                          AnglesVec=angles, ObjSize=detX, device_projector=0, OS_number=os_n)


def _dir(angles, detX, detY, pad=0):
    from tomobar_b200.methodsDIR_CuPy import RecToolsDIRCuPy

    return RecToolsDIRCuPy(DetectorsDimH=detX, DetectorsDimH_pad=pad, DetectorsDimV=detY, CenterRotOffset=0.0,
                           AnglesVec=angles, ObjSize=detX, device_projector=0)


LABELS = ["angles", "detY", "detX"]


def test_forwproj_ones(gpu_scan):  # tests/test_RecToolsDIRCuPy.py:669-695
    data, angles = gpu_scan
    R = _dir(angles, 160, 128)
    fp = R.FORWPROJ(torch.ones(128, 160, 160, device="cuda")).cpu().numpy()
    assert_allclose(fp.min(), 67.27458, rtol=2e-6)
    assert_allclose(fp.max(), 225.27428, rtol=2e-6)
    assert fp.dtype == np.float32 and fp.shape == (128, 180, 160)


def test_backproj_view_bug_compat(gpu_scan):  # tests/test_RecToolsDIRCuPy.py:695-717 (the CuPy path's own golden)
    """The reference back-projects the raw buffer of the swapped view; compat_view_bug=True reproduces its golden."""
    data, angles = gpu_scan
    R = _dir(angles, 160, 128)
    R.compat_view_bug = True
    bp = R.BACKPROJ(data, data_axes_labels_order=LABELS).cpu().numpy()
    assert_allclose(bp.max(), 174.80643, rtol=2e-6)
    assert_allclose(bp.min(), -2.309583, rtol=2e-4)  # the minimum sits on a scrambled edge: 1e-4 from the golden
    assert bp.shape == (128, 160, 160)


def test_backproj(gpu_scan):  # tests/test_RecToolsDIR.py:221-240 (host-array path, correct layout)
    data, angles = gpu_scan
    R = _dir(angles, 160, 128)
    bp = R.BACKPROJ(data, data_axes_labels_order=LABELS).cpu().numpy()
    assert_allclose(bp.min(), -3.8901403, rtol=1e-6)
    assert_allclose(bp.max(), 350.38193, rtol=1e-6)
    assert bp.shape == (128, 160, 160)


def test_fbp3d(gpu_scan):  # tests/test_RecToolsDIRCuPy.py:543-566
    data, angles = gpu_scan
    R = _dir(angles, 160, 128)
    rec = R.FBP(data, data_axes_labels_order=LABELS, cutoff_freq=1.1).cpu().numpy()
    assert_allclose(rec.min(), -0.014693323, rtol=2e-6)
    assert_allclose(rec.max(), 0.0340156, rtol=2e-6)
    assert rec.dtype == np.float32 and rec.shape == (128, 160, 160)


def test_fbp3d_pad(gpu_scan):  # :569-591
    data, angles = gpu_scan
    R = _dir(angles, 160, 128, pad=20)
    rec = R.FBP(data, data_axes_labels_order=LABELS, cutoff_freq=1.1).cpu().numpy()
    assert_allclose(rec.min(), -0.013320832, rtol=1e-5)
    assert_allclose(rec.max(), 0.03534874, rtol=1e-5)
    assert rec.shape == (128, 160, 160)


def test_fbp3d_mask(gpu_scan):  # :644-667
    data, angles = gpu_scan
    R = _dir(angles, 160, 128)
    rec = R.FBP(data, data_axes_labels_order=LABELS, recon_mask_radius=0.7, cutoff_freq=1.1).cpu().numpy()
    assert_allclose(rec.min(), -0.0129751, rtol=2e-6)
    assert_allclose(rec.max(), 0.0340156, rtol=2e-6)


def test_landweber_3d(gpu_scan):  # tests/test_RecToolsIRCuPy.py:12-40
    data, angles = gpu_scan
    rec = _ir(angles, 160, 128).Landweber({"projection_data": data, "data_axes_labels_order": LABELS},
                                          {"iterations": 10}).cpu().numpy()
    assert_allclose(rec.min(), -0.00026702078, rtol=1e-6)
    assert_allclose(rec.max(), 0.016753351, rtol=1e-6)
    assert rec.dtype == np.float32 and rec.shape == (128, 160, 160)


def test_landweber_2d(gpu_scan):  # :43-69 (200 iterations; oracle drift 2.4e-5)
    data, angles = gpu_scan
    from tomobar_b200.methodsIR_CuPy import RecToolsIRCuPy

    R = RecToolsIRCuPy(160, 0, None, 0.0, angles, 160, 0, None)
    rec = R.Landweber({"projection_data": data[:, 64, :], "data_axes_labels_order": ["angles", "detX"]},
                      {"iterations": 200})
    assert rec.shape == (1, 160, 160)
    rec = rec[0].cpu().numpy()
    assert_allclose(rec.min(), -0.0027037817, rtol=1e-4)
    assert_allclose(rec.max(), 0.02463191, rtol=1e-4)


def test_sirt_3d(gpu_scan):  # :98-126 (oracle: 1.6e-4 on the min)
    data, angles = gpu_scan
    rec = _ir(angles, 160, 128).SIRT({"projection_data": data, "data_axes_labels_order": LABELS},
                                     {"iterations": 5}).cpu().numpy()
    assert_allclose(rec.min(), -0.0011388711, rtol=5e-4)
    assert_allclose(rec.max(), 0.020178854, rtol=1e-4)


def test_powermethod(gpu_scan):  # :224-247, :250-272
    data, angles = gpu_scan
    lc = _ir(angles, 160, 128).powermethod({"projection_data": data, "data_axes_labels_order": LABELS})
    assert 27200 <= lc <= 27800
    assert_allclose(lc, 27550.463, rtol=1e-4)  # value the reference's FISTA tests hard-code
    lc_os = _ir(angles, 160, 128, os_n=5).powermethod({"projection_data": data, "data_axes_labels_order": LABELS})
    assert 5200 <= lc_os <= 5700
    assert_allclose(lc_os, 5510.867, rtol=1e-4)


def test_fista_3d(gpu_scan):  # :297-323
    data, angles = gpu_scan
    rec = _ir(angles, 160, 128).FISTA({"projection_data": data, "data_axes_labels_order": LABELS},
                                      {"iterations": 10, "lipschitz_const": 27550.463}).cpu().numpy()
    # the golden is quoted to 3 digits (-0.00214); the restated ASTRA model gives -0.00214049
    # (2.3e-4 from it), the max agrees to 2e-5
    assert_allclose(rec.min(), -0.00214, rtol=3e-4)
    assert_allclose(rec.max(), 0.024637, rtol=1e-4)
    assert rec.dtype == np.float32 and rec.shape == (128, 160, 160)


def test_fbp3d_swapped_axes_pad(gpu_scan):  # tests/test_RecToolsDIRCuPy.py:593-621
    data, angles = gpu_scan
    swapped = data.swapaxes(0, 1).swapaxes(0, 2)  # ["detX", "angles", "detY"]
    assert tuple(swapped.shape) == (160, 180, 128)
    R = _dir(angles, 160, 128, pad=45)
    rec = R.FBP(swapped, data_axes_labels_order=["detX", "angles", "detY"], cutoff_freq=1.1).cpu().numpy()
    assert_allclose(rec.mean(), 0.001496, atol=1e-4)
    assert rec.dtype == np.float32 and rec.shape == (128, 160, 160)


def test_fbp3d_from_raw_data():  # tests/test_RecToolsDIR.py:305-323 (normaliser + FBP, sinc cutoff 1.1)
    import golden_cases as G
    from tomobar_b200.supp.suppTools import normaliser

    scan = G.load_scan()
    if "raw" not in scan:
        pytest.skip("tests/golden/tomo_standard.npz missing")
    normalised = normaliser(*scan["raw"])
    R = _dir(scan["angles"], 160, 128)
    rec = R.FBP(normalised, data_axes_labels_order=LABELS, cutoff_freq=1.1).cpu().numpy()
    assert_allclose(rec.min(), -0.014656051, rtol=1e-5)
    assert_allclose(rec.max(), 0.0338298, rtol=1e-5)
    assert rec.dtype == np.float32 and rec.shape == (128, 160, 160)


def test_sirt_x3(gpu_scan):  # the SIRT half of tests/test_RecToolsIRCuPy.py:190-222
    data, angles = gpu_scan
    rec = _ir(angles, 160, 128).SIRT({"projection_data": data, "data_axes_labels_order": LABELS},
                                     {"iterations": 3}).cpu().numpy()
    # R = 1/(A 1) amplifies grazing-ray differences at the volume corners, where the minimum sits:
    # the restated ASTRA model gives -0.00028118 (1.7e-3 from the golden)
    assert_allclose(rec.min(), -0.0002806916, rtol=3e-3)
    assert rec.shape == (128, 160, 160)


def test_powermethod_os_pwls_pad(gpu_scan):  # :273-295
    data, angles = gpu_scan
    lc = _ir(angles, 160, 128, pad=50, os_n=5).powermethod(
        {"data_fidelity": "PWLS", "projection_data": data, "data_axes_labels_order": LABELS})
    assert 8000 <= lc <= 9000
