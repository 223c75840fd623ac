import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running oracle checks")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def scan():
    """The reference's own test scan (tests/conftest.py:110-121 there): data_norm
    (180 angles, 128 detY, 160 detX) float32, angles (180,) float32."""
    d = np.load(os.path.join(GOLDEN, "normalised_data.npz"))
    return d["data_norm"], d["angles"]


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.build()
    return O


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def rel_max(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


import contextlib


@contextlib.contextmanager
def single_iteration_tv():
    """Whole-volume PD_TV through the one-iteration-per-launch strip kernel: the z-sharded prox launches
    exactly that kernel, so sharded == unsharded is a bit-for-bit statement about it (the default
    unsharded prox pairs iterations in the fused kernel, same arithmetic, equal to ~1e-7)."""
    from tomobar_b200._lib import lib

    old = lib.tmb_tv_set_simple_kernels(3)
    try:
        yield
    finally:
        lib.tmb_tv_set_simple_kernels(old)
