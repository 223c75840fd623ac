"""FBP sinc filtering on the GPU: cuFFT (through torch.fft) around two libtmb kernels
(behaviour of tomobar/fourier.py:26-78 and cuda_kernels/generate_filtersync.cu)."""

from __future__ import annotations

import torch

from tomobar_b200._lib import lib, check
from tomobar_b200._tensors import ptr, stream_ptr


def sinc_filter(n: int, cutoff: float, multiplier: float, device) -> torch.Tensor:
    """Half-spectrum sinc-ramp filter f[n//2+1] (generate_filtersync.cu:5-82)."""
    f = torch.empty(n // 2 + 1, dtype=torch.float32, device=device)
    with torch.cuda.device(f.device):
        check(lib.tmb_sinc_filter(float(cutoff), ptr(f), int(n), float(multiplier), stream_ptr(f)),
              "tmb_sinc_filter")
    return f


def _filtersinc3D_cupy(projection3D: torch.Tensor, cutoff: float = 0.6) -> torch.Tensor:
    """irfft(rfft(p) * f) along detX for p[angles, detY, detX]; the 1/(angles*detX) scaling is
    folded into the filter (fourier.py:52-71)."""
    projectionsNum, _, DetectorsLengthH = projection3D.shape
    proj_f = torch.fft.rfft(projection3D, dim=-1, norm="backward").contiguous()
    f = sinc_filter(DetectorsLengthH, cutoff, 1.0 / projectionsNum / DetectorsLengthH, projection3D.device)
    rows = proj_f.numel() // proj_f.shape[-1]
    spec = torch.view_as_real(proj_f)
    with torch.cuda.device(spec.device):
        check(lib.tmb_apply_filter(ptr(spec), ptr(f), rows, proj_f.shape[-1], stream_ptr(spec)),
              "tmb_apply_filter")
    return torch.fft.irfft(proj_f, n=DetectorsLengthH, dim=-1, norm="forward")


# ----------------------------------------------------------------------------------------------
# analytic filters of the Fourier (USFFT) reconstruction -- host side, numpy
# (behaviour of tomobar/fourier.py:81-159, after V. Nikitin's tomocupy)
# ----------------------------------------------------------------------------------------------
import functools  # noqa: E402

import numpy as np  # noqa: E402


def _wint(order: int, t: np.ndarray) -> np.ndarray:
    """Quadrature weights for  int t f(t) dt  on the grid ``t`` from piecewise polynomials of
    degree ``order - 1`` through ``order`` consecutive nodes, overlapping windows averaged
    (fourier.py:81-108).  The last 40 weights are replaced by a linear ramp."""
    count = len(t)
    nodes = np.linspace(1e-40, 1, order)
    log_nodes = np.log(nodes)
    powers = np.arange(order)
    # monomial basis evaluated at the unit-interval nodes, inverted
    inv_vandermonde = np.linalg.inv(np.exp(np.outer(powers, log_nodes)))
    k = np.arange(1, order + 2)
    # integrals of the monomials x^(k-1) between consecutive nodes
    seg = np.diff(np.exp(np.outer(k, log_nodes)) * np.tile(1.0 / k[..., np.newaxis], [1, order]))
    lin_term = np.matmul(inv_vandermonde, seg[1:order + 1, :])   # x * p(x)
    const_term = np.matmul(inv_vandermonde, seg[0:order, :])     # const * p(x)
    # each short interval is covered by up to (order - 1) windows
    cover = 1 / np.concatenate((np.arange(1, order), (order - 1) * np.ones((count - 2 * (order - 1) - 1)),
                                np.arange(order - 1, 0, -1)))
    w = np.zeros(count)
    for j in range(count - order + 1):
        span = t[j + order - 1] - t[j]
        local = (span ** 2) * lin_term + span * t[j] * const_term
        w[j:j + order] += local @ cover[j:j + order - 1]
    w[-40:] = (w[-40]) / (count - 40) * np.arange(count - 40, count)
    return w


_WINDOWS = {
    "ramp": lambda t, d: 1.0,
    "shepp": lambda t, d: np.sinc(t / (2 * d)) * (t / d <= 2),
    "cosine": lambda t, d: np.cos(np.pi * t / (2 * d)) * (t / d <= 1),
    "cosine2": lambda t, d: (np.cos(np.pi * t / (2 * d))) ** 2 * (t / d <= 1),
    "hamming": lambda t, d: (0.54 + 0.46 * np.cos(np.pi * t / d)) * (t / d <= 1),
    "hann": lambda t, d: (1 + np.cos(np.pi * t / d)) / 2.0 * (t / d <= 1),
    "parzen": lambda t, d: pow(1 - t / d, 3) * (t / d <= 1),
}


def calc_filter(n: int, filter: str, cutoff_freq: float) -> np.ndarray:
    """Half-spectrum FBP filter (n // 2 + 1 bins, float32) for the Fourier reconstruction
    (fourier.py:111-159).  The quadrature weights are a host loop over the bins (13 ms at 8192 points, a seventh of a
    config-4 FOURIER_INV call during which the device idles): the table is computed once per (n, filter, cutoff)."""
    return _calc_filter_table(int(n), str(filter), float(cutoff_freq)).copy()


@functools.lru_cache(maxsize=32)
def _calc_filter_table(n: int, filter: str, cutoff_freq: float) -> np.ndarray:
    d = 0.5
    t = np.arange(0, n / 2 + 1) / n
    if filter == "none":
        return np.asarray(n * cutoff_freq + t * 0, dtype=np.float32)
    if filter not in _WINDOWS:
        raise ValueError(f"unknown filter {filter!r}")
    wfa = n * cutoff_freq * _wint(12, t) * _WINDOWS[filter](t, d)
    wfa = 2 * wfa * (wfa >= 0)
    wfa[0] *= 2
    return np.asarray(wfa, dtype=np.float32)
