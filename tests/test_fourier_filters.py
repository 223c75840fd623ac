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
