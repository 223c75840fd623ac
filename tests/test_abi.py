"""CPU-side checks of the C-ABI boundary: libtmb.so loads without a GPU, exports every symbol
include/tmb.h declares, and its host-side geometry matches the oracle's."""

import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "tmb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tmb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from tomobar_b200._lib import lib, SIGNATURES

    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in tmb.h but not exported by libtmb.so"
    assert set(SIGNATURES) == set(names), "python binding table and tmb.h disagree"


def test_geometry_table_matches_oracle(oracle):
    from tomobar_b200.projector import ProjTools3D

    rng = np.random.default_rng(0)
    for dtype in (np.float32, np.float64):
        angles = np.linspace(0, np.pi, 37, endpoint=False).astype(dtype)
        cor = 1.75
        P = ProjTools3D(50, 3, 5, angles, cor, 48, "gpu", 0, None)
        tbl = P.angle_table()
        ref = oracle.angle_table(angles, cor, 48, 56)
        np.testing.assert_array_equal(tbl, ref)
    cor_vec = rng.uniform(-2, 2, 37)
    P = ProjTools3D(50, 0, 5, angles, cor_vec, 50, "gpu", 0, None)
    np.testing.assert_array_equal(P.angle_table(), oracle.angle_table(angles, cor_vec, 50, 50))


def test_subset_table_matches_reference_format(oracle):
    from tomobar_b200.projector import ProjTools3D

    for na, os_n in [(180, 5), (180, 7), (37, 6), (12, 12), (1800, 24)]:
        angles = np.linspace(0, np.pi, na, endpoint=False).astype(np.float32)
        P = ProjTools3D(16, 0, 2, angles, 0.0, 16, "gpu", 0, os_n)
        tab, bins = oracle.os_indices(na, os_n)
        assert P.NumbProjBins == bins
        np.testing.assert_array_equal(P.newInd_Vec, tab)
        for s in range(os_n):
            assert P.subset_size(s) == len(range(s, na, os_n))


def test_bad_arguments_raise():
    import pytest
    from tomobar_b200.projector import ProjTools3D

    a = np.zeros(4, np.float32)
    with pytest.raises(ValueError):
        ProjTools3D(0, 0, 1, a, 0.0, 8)
    with pytest.raises(ValueError):
        ProjTools3D(8, -1, 1, a, 0.0, 8)
    with pytest.raises(ValueError):
        ProjTools3D(8, 0, 1, np.zeros((2, 2), np.float32), 0.0, 8)
    with pytest.raises(ValueError):
        ProjTools3D(8, 0, 1, a, np.zeros(3), 8)
    with pytest.raises(ValueError):
        ProjTools3D(8, 0, 1, a, 0.0, (8, 8))
    with pytest.raises(ValueError):
        ProjTools3D(8, 0, 1, a, 0.0, 8, "cpu")
    with pytest.raises(ValueError):
        ProjTools3D(8, 0, 1, a, 0.0, 8, "gpu", 0, 0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tomobar_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "__never__", f"{f} mentions the oracle"


def test_pd_tv_launch_accounting_needs_no_gpu():
    """tmb_pd_tv_launches is host arithmetic: pairs of iterations share a launch where the fused
    kernel applies (fp32 duals, 3-D, rows of whole float4s), every iteration launches otherwise."""
    from tomobar_b200._lib import lib

    assert lib.tmb_tv_set_simple_kernels(0) in range(0, 11)
    try:
        assert lib.tmb_pd_tv_launches(512, 2048, 2048, 50, 0) == 25
        assert lib.tmb_pd_tv_launches(512, 2048, 2048, 7, 0) == 4      # odd tail: one single iteration
        assert lib.tmb_pd_tv_launches(512, 2048, 2048, 50, 1) == 50     # fp16 duals: strip kernel
        assert lib.tmb_pd_tv_launches(16, 64, 150, 10, 0) == 10         # dx % 4 != 0
        assert lib.tmb_pd_tv_launches(1, 64, 64, 10, 0) == 10           # 2-D
        assert lib.tmb_pd_tv_launches(16, 64, 64, 0, 0) == 0
        for mode, want in ((1, 10), (2, 10), (3, 10), (4, 10), (5, 5), (6, 5), (7, 5), (8, 5), (9, 5), (10, 5)):
            lib.tmb_tv_set_simple_kernels(mode)
            assert lib.tmb_pd_tv_launches(16, 64, 64, 10, 0) == want, mode
    finally:
        lib.tmb_tv_set_simple_kernels(0)


def test_forward_projector_group_size_host_logic():
    """Which forward-projector family a geometry gets (host arithmetic of tmb_geom_fp_group, no GPU): the
    multi-angle kernel k_fpm with groups of 4 at BASELINE.json's sizes, smaller groups / one angle per CTA
    (0) where neighbouring subset angles diverge by more than its staged window holds, and stacks of at
    most 16 slices or volumes of a single line segment stay with k_fp / k_fpq."""
    from tomobar_b200._lib import lib
    from tomobar_b200.projector import ProjTools3D

    def group(nz, n, na, os_n, sub=0):
        angles = np.linspace(0, np.pi, na, endpoint=False).astype(np.float32)
        P = ProjTools3D(n, 0, nz, angles, 0.0, n, "gpu", 0, os_n)
        return lib.tmb_geom_fp_group(P._g, sub)

    assert group(512, 2048, 1800, 24) == 4      # headline
    assert group(64, 2048, 1800, 24, 23) == 4   # a z-shard of it
    assert group(256, 1024, 900, 6) == 4        # config 2
    assert group(384, 1536, 1500, 6) == 4       # config 5
    assert group(8, 2048, 1800, 24) == 0        # <= 16 slices: k_fp
    assert group(64, 256, 180, 1) == 0          # one line segment: k_fpq writes the sinogram itself
    coarse = group(64, 2048, 96, 8)             # subset angles 15 degrees apart: windows too wide for any group
    assert coarse == 0
    mid = group(64, 2048, 300, 12)              # 7.2 degrees apart: smaller groups
    assert 2 <= mid < 4
