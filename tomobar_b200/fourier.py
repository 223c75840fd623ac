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
