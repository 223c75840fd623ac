"""calc_filter medians pinned by the reference (tests/test_fourier.py:6-26); CPU only."""

import numpy as np
import pytest
from numpy.testing import assert_allclose


@pytest.mark.parametrize("name,median", [("none", 100), ("ramp", 0.496701), ("shepp", 0.447188),
                                         ("cosine", 0.25168), ("cosine2", 0.164889), ("hamming", 0.185245),
                                         ("hann", 0.164889), ("parzen", 0.042508)])
def test_calc_filter(name, median):
    from tomobar_b200.fourier import calc_filter

    f = calc_filter(100, name, 1.0)
    assert f.size == 100 / 2 + 1 and f.dtype == np.float32
    f = np.sort(f)
    assert_allclose(f[f.size // 2], median, rtol=1e-5)


def test_calc_filter_table_is_cached_and_callers_get_their_own_copy():
    """The quadrature-weight loop of calc_filter (fourier.py:81-159 of the reference) runs once per (n, filter, cutoff);
    every caller gets a private, writable copy of the cached table."""
    from tomobar_b200.fourier import _calc_filter_table, calc_filter

    _calc_filter_table.cache_clear()
    a = calc_filter(512, "shepp", 1.0)
    hits0 = _calc_filter_table.cache_info().hits
    b = calc_filter(512, "shepp", 1.0)
    assert _calc_filter_table.cache_info().hits == hits0 + 1
    assert a is not b and a.flags.writeable and np.array_equal(a, b)
    a[:] = -1.0
    np.testing.assert_array_equal(calc_filter(512, "shepp", 1.0), b)
    assert not np.array_equal(calc_filter(512, "hann", 1.0), b)
    assert not np.array_equal(calc_filter(512, "shepp", 0.5), b)
