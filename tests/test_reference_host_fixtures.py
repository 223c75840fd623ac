"""Host logic and oracle pinned to outputs of the reference's OWN pure-numpy functions.

``tests/golden/reference_host_fixtures.npz`` is written by ``tests/golden/make_fixtures.py``
(``reference_host_fixtures``), which imports supp/suppTools.py, supp/funcs.py and fourier.py from
/root/reference file by file and stores what they return.  CPU tests pin the oracle and the host
side; the ``gpu`` tests pin the two CUDA kernels behind ``normaliser`` / ``apply_circular_mask``."""

import ast
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(HERE, "golden", "reference_host_fixtures.npz"))


@pytest.fixture(scope="module")
def raw():
    return np.load(os.path.join(HERE, "golden", "tomo_standard.npz"))


FILTERS = ("none", "ramp", "shepp", "cosine", "cosine2", "hamming", "hann", "parzen")


# ---------------------------------------------------------------- CPU: host logic and oracle
@pytest.mark.parametrize("name", FILTERS)
def test_calc_filter_matches_reference(ref, name):
    """fourier.py:122-166 (calc_filter with its _wint integration weights)."""
    from tomobar_b200.fourier import calc_filter

    for n in (100, 128, 4096):
        for cut in (1.0, 0.35):
            want = ref[f"filt_{name}_{n}_{cut}"]
            got = calc_filter(n, name, cut)
            assert got.dtype == np.float32 and got.shape == want.shape
            assert_allclose(got, want, rtol=2e-6, atol=1e-7 * np.abs(want).max())


def test_axis_permutations_match_reference_swaps(ref):
    """funcs.py:99-141,190-206: the reference's swap list applied to a shape tuple must land where
    our single permutation lands, for all orders of the 3-D and 2-D labels."""
    from tomobar_b200.supp.funcs import _data_dims_swapper

    want3, want2 = ["detY", "angles", "detX"], ["angles", "detX"]
    sizes = {"angles": 7, "detX": 11, "detY": 5}
    for row in ref["axis_swaps"]:
        labels, swaps = str(row).split("=")
        labels = labels.split("|")
        shape = [sizes[l] for l in labels]
        for sw in ast.literal_eval(swaps):
            if sw is not None:
                shape[sw[0]], shape[sw[1]] = shape[sw[1]], shape[sw[0]]
        required = want3 if len(labels) == 3 else want2
        assert tuple(shape) == tuple(sizes[l] for l in required)
        assert _data_dims_swapper(tuple(sizes[l] for l in labels), labels, required) == tuple(shape)


def test_axis_permutation_of_tensors(ref):
    import torch
    from tomobar_b200.supp.funcs import _data_dims_swapper

    sizes = {"angles": 7, "detX": 11, "detY": 5}
    for row in ref["axis_swaps"]:
        labels, swaps = str(row).split("=")
        labels = labels.split("|")
        required = ["detY", "angles", "detX"] if len(labels) == 3 else ["angles", "detX"]
        a = np.arange(np.prod([sizes[l] for l in labels]), dtype=np.float32).reshape([sizes[l] for l in labels])
        want = a
        for sw in ast.literal_eval(swaps):
            if sw is not None:
                want = np.swapaxes(want, sw[0], sw[1])
        got = _data_dims_swapper(torch.from_numpy(a), labels, required)
        assert_array_equal(got.numpy(), want)


@pytest.mark.parametrize("key_a,key_v,cor", [("geom_angles32", "geom_vec32_cor0", 0.0),
                                             ("geom_angles64", "geom_vec64_cor", 3.25)])
def test_angle_table_matches_reference_vectors(ref, key_a, key_v, cor):
    """funcs.py:45-65: ray = R(theta)(0,-1,0), detector centre = R(theta)(cor,0,0),
    u = R(theta)(1,0,0), v = (0,0,1).  The oracle's table must carry exactly the reference's
    cos/sin (evaluated in the dtype of the angles) and its detector shift."""
    from oracle import oracle as orc

    ang, vec = ref[key_a], ref[key_v]
    n, nu = 32, 48
    tbl = orc.angle_table(ang, cor, n, nu)
    assert_array_equal(tbl[:, 0], vec[:, 6].astype(np.float32))      # cos = u_x
    assert_array_equal(tbl[:, 1], vec[:, 7].astype(np.float32))      # sin = u_y
    assert_array_equal(vec[:, 0].astype(np.float32), tbl[:, 1])      # ray_x = sin
    assert_array_equal(vec[:, 1].astype(np.float32), -tbl[:, 0])     # ray_y = -cos
    assert_array_equal(vec[:, [2, 5, 8, 9, 10]], 0.0)
    assert_array_equal(vec[:, 11], 1.0)
    # detector centre projected on u is the centre-of-rotation offset the table stores
    shift = vec[:, 3] * vec[:, 6] + vec[:, 4] * vec[:, 7]
    assert_allclose(tbl[:, 2], -shift + (nu / 2.0 - 0.5), rtol=1e-6)


def test_oracle_circular_mask_matches_reference(ref):
    """suppTools.py:364-396."""
    from oracle import oracle as orc

    for i, (n, rad) in enumerate(ref["mask_cases"]):
        n = int(n)
        want = np.unpackbits(ref[f"mask_{i}"])[: n * n].reshape(n, n).astype(bool)
        got = orc.circular_mask(np.ones((n, n), np.float32), float(rad)) > 0
        assert_array_equal(got, want, err_msg=f"n={n} radius={rad}")
        vol = orc.circular_mask(np.ones((3, n, n), np.float32), float(rad)) > 0
        assert_array_equal(vol, np.broadcast_to(want, (3, n, n)))


def test_golden_case_normaliser_matches_reference(ref, raw):
    """suppTools.py:187-264 (mean): the numpy restatement the golden cases are fed through."""
    from golden_cases import normaliser_mean

    # float32 inputs, like the reference's own fixtures (tests/conftest.py:93-106)
    norm = normaliser_mean(*(np.float32(raw[k]) for k in ("data", "flats", "darks")))
    assert_array_equal(norm[::9, ::8, ::8], ref["norm_mean_sub"])
    mn, mx, mean = ref["norm_mean_stats"]
    assert norm.min() == mn and norm.max() == mx
    assert_allclose(norm.mean(dtype=np.float64), mean, rtol=1e-12)


# ---------------------------------------------------------------- GPU: the two CUDA kernels
@pytest.mark.gpu
def test_cuda_circular_mask_matches_reference(ref):
    import torch
    from tomobar_b200.supp.suppTools import apply_circular_mask

    for i, (n, rad) in enumerate(ref["mask_cases"]):
        n = int(n)
        want = np.unpackbits(ref[f"mask_{i}"])[: n * n].reshape(n, n).astype(bool)
        for shape in ((n, n), (3, n, n)):
            x = torch.ones(shape, device="cuda")
            got = apply_circular_mask(x, float(rad)).cpu().numpy() > 0
            assert_array_equal(got, np.broadcast_to(want, shape), err_msg=f"n={n} radius={rad}")


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["mean", "median"])
def test_cuda_normaliser_matches_reference(ref, raw, method):
    """fp32 tolerance: 1e-5 relative on -log(...) (the reference divides in float32 with numpy,
    the kernel with IEEE division and logf)."""
    from tomobar_b200.supp.suppTools import normaliser

    out = normaliser(raw["data"], raw["flats"], raw["darks"], method=method).cpu().numpy()
    want = ref[f"norm_{method}_sub"]
    assert_allclose(out[::9, ::8, ::8], want, rtol=1e-5, atol=2e-6)
    mn, mx, mean = ref[f"norm_{method}_stats"]
    assert_allclose([out.min(), out.max(), out.mean(dtype=np.float64)], [mn, mx, mean], rtol=1e-5, atol=2e-6)


@pytest.mark.gpu
def test_cuda_normaliser_without_log_matches_reference(ref, raw):
    from tomobar_b200.supp.suppTools import normaliser

    out = normaliser(raw["data"], raw["flats"], raw["darks"], log=False).cpu().numpy()
    assert_allclose(out[::9, ::8, ::8], ref["norm_nolog_sub"], rtol=3e-6, atol=1e-7)  # means summed in another order
