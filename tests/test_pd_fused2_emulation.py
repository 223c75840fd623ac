"""The state machine of the fused two-iteration PD_TV kernel (k_pd_tv3d_f2), replayed on the CPU by
tools/emulate_pd_fused2.py: every voxel stored exactly once and equal, bit for bit, to two plain
iterations -- at window, strip, z-run and volume edges.  (The CUDA kernel itself: tests/test_gpu_tv.py.)"""

import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu():
    spec = importlib.util.spec_from_file_location("emulate_pd_fused2", os.path.join(ROOT, "tools", "emulate_pd_fused2.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("shape,zrun,nonneg,aniso", [
    ((2, 3, 8), 2, False, False),        # smallest volume: every plane is a boundary plane
    ((5, 9, 124), 5, True, False),       # two windows, the second holds 4 columns
    ((7, 18, 132), 3, False, False),     # three z-runs, strips cut by the last row
    ((9, 21, 244), 4, True, True),       # anisotropic projection
    ((6, 16, 120), 2, False, False),     # exactly one window / one CTA row
    ((4, 5, 4), 1, False, False),        # one column group: first and last column in the same lane
    ((3, 2, 12), 3, False, False),
])
def test_emulated_kernel_equals_two_plain_iterations(emu, shape, zrun, nonneg, aniso):
    assert emu.run_case(shape, zrun, nonneg, aniso, seed=sum(shape))


@pytest.mark.parametrize("shape,cuts,zrun,nonneg,aniso", [
    ((8, 9, 124), [4], 4, True, False),
    ((9, 6, 12), [2, 5], 2, False, False),      # shards of 2, 3 and 4 planes
    ((10, 18, 132), [3, 7], 8, False, True),
    ((6, 5, 8), [2, 4], 1, True, False),        # every z-run starts in the neighbour's planes
])
def test_emulated_sharded_kernel_equals_two_plain_iterations(emu, shape, cuts, zrun, nonneg, aniso):
    """The GHOST variant (tmb_pd_tv_iter2): every z-shard emulated on its own, reading two ghost planes
    of U and one of P / Input from its neighbours' arrays; assembled result == whole volume, bit for bit."""
    assert emu.run_sharded_case(shape, cuts, zrun, nonneg, aniso, seed=sum(shape) + len(cuts))
