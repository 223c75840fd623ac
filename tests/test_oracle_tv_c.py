"""The OpenMP twins of the oracle's TV operators (oracle/tv_oracle.c, used so that bench.py's CPU baseline runs the
whole sub-step on all host cores) are BIT-IDENTICAL to the numpy restatements of the reference kernels."""

import numpy as np
import pytest

from oracle import oracle as O


@pytest.mark.parametrize("shape", [(7, 12, 20), (1, 16, 24), (16, 1, 24), (5, 9, 8), (2, 2, 2), (33, 17, 40)])
@pytest.mark.parametrize("methodTV,nonneg", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_pd_tv_c_equals_numpy(shape, methodTV, nonneg):
    rng = np.random.default_rng(sum(shape))
    v = (rng.standard_normal(shape) * 0.05 + (rng.random(shape) > 0.6) * 0.1).astype(np.float32)
    for lam, lip in ((4e-4, 12.0), (3e-2, 8.0)):
        a = O.pd_tv(v, lam, 11, methodTV, nonneg, lip, False, use_c=True)
        b = O.pd_tv(v, lam, 11, methodTV, nonneg, lip, False, use_c=False)
        assert a.shape == b.shape and np.array_equal(a, b)


@pytest.mark.parametrize("shape", [(7, 12, 20), (1, 16, 24), (16, 1, 24), (5, 9, 8), (2, 2, 2), (33, 17, 40)])
def test_rof_tv_c_equals_numpy(shape):
    rng = np.random.default_rng(sum(shape) + 1)
    v = (rng.standard_normal(shape) * 0.05 + (rng.random(shape) > 0.6) * 0.1).astype(np.float32)
    for lam, tau in ((4e-4, 1e-3), (2e-2, 5e-3)):
        a = O.rof_tv(v, lam, 11, tau, False, use_c=True)
        b = O.rof_tv(v, lam, 11, tau, False, use_c=False)
        assert a.shape == b.shape and np.array_equal(a, b)


def test_thread_count_is_reported():
    assert O.threads() >= 1
