"""Dry-run memory accounting (the reference's DeviceMemStack protocol, memory_estimator_helpers.py and
methodsDIR_CuPy.py:253-258, 437-441): host arithmetic only, no GPU."""

import numpy as np
import pytest


def _rec(recon_size, pad=0):
    from tomobar_b200.methodsDIR_CuPy import RecToolsDIRCuPy

    r = object.__new__(RecToolsDIRCuPy)  # the estimator needs no projector (and no GPU)
    r.recon_size, r.detectors_x_pad = recon_size, pad
    return r


def test_device_mem_stack_semantics():
    from tomobar_b200.supp.memory_estimator_helpers import DeviceMemStack

    assert DeviceMemStack.instance() is None
    with DeviceMemStack() as outer:
        assert DeviceMemStack.instance() is outer
        outer.malloc(1)
        outer.malloc(513)
        assert outer.current == 512 + 1024 and outer.highwater == 1536 and outer.allocations == [1, 513]
        with DeviceMemStack() as inner:                 # nested blocks report to the outermost stack
            assert DeviceMemStack.instance() is outer and inner.highwater == 0
        outer.free(513)
        assert outer.current == 512 and outer.highwater == 1536
        with pytest.raises(AssertionError):
            outer.free(7)
    assert DeviceMemStack.instance() is None


def test_fourier_inv_estimator_config4():
    """BASELINE.json config 4 (128 x 2000 x 2048): the peak is the inverse 2-D FFT of a chunk of complex slices -- the
    caller's input, the polar samples, the reconstruction, and the chunk's grid + transform + work area (the oversampled
    grid never exists for the whole volume)."""
    from tomobar_b200.supp.memory_estimator_helpers import DeviceMemStack

    nz, nproj, n = 128, 2000, 2048
    with DeviceMemStack() as st:
        shape = _rec(n).FOURIER_INV((nz, nproj, n), data_dtype=np.float32)
    assert shape == (nz, n, n)
    inp, recon = nz * nproj * n * 4, nz * n * n * 4
    chunk = (1 << 28) // (4 * n * n)
    piece = chunk * (2 * n) ** 2 * 8
    small = 3 * 8192                                     # theta, sorted theta, indices (rounded to 512 B)
    assert st.highwater == 2 * inp + recon + 3 * piece + small
    assert st.current == inp                             # only the caller's array is left
    assert 12.7e9 < st.highwater < 12.9e9


def test_fourier_inv_estimator_options():
    from tomobar_b200.supp.memory_estimator_helpers import DeviceMemStack

    def peak(shape, rec=None, **kw):
        with DeviceMemStack() as st:
            out = (rec or _rec(shape[2])).FOURIER_INV(shape, **kw)
        return st.highwater, out

    base, out = peak((64, 900, 1024))
    assert out == (64, 1024, 1024)
    # odd sizes are padded to even ones (an extra copy of the projections during the filter step)
    odd, out_odd = peak((63, 900, 1023), _rec(1023))
    assert out_odd == (63, 1023, 1023) and odd > 0
    # more slices, more memory; axis labels are honoured; detector padding widens the grid
    assert peak((128, 900, 1024))[0] > base
    assert peak((900, 64, 1024), data_axes_labels_order=["angles", "detY", "detX"])[0] == base
    assert peak((64, 900, 1024), _rec(1024, pad=128))[0] > base
    assert peak((64, 900, 1024), padding=64)[0] > base
    # the crop of the reconstruction
    assert peak((64, 900, 1024), _rec(512))[1] == (64, 512, 512)
    with pytest.raises(ValueError):
        peak((64, 900, 1024), _rec(2048))
    with pytest.raises(ValueError):                      # a shape tuple without an active stack
        _rec(1024).FOURIER_INV((64, 900, 1024))
