"""Detector padding, reconstruction cropping and circular masking on CUDA tensors
(behaviour of tomobar/supp/suppTools.py:364-467)."""

from __future__ import annotations

import torch

from tomobar_b200._lib import lib, check
from tomobar_b200._tensors import ptr, stream_ptr


def _apply_horiz_detector_padding(data: torch.Tensor, detector_width_pad: int, cupyrun: bool = True) -> torch.Tensor:
    """Edge-pad detX (the last axis) on both sides (suppTools.py:425-459)."""
    if detector_width_pad <= 0:
        return data
    p = int(detector_width_pad)
    left = data[..., :1].expand(*data.shape[:-1], p)
    right = data[..., -1:].expand(*data.shape[:-1], p)
    return torch.cat((left, data, right), dim=-1)


def perform_recon_crop(data: torch.Tensor, croped_size: int) -> torch.Tensor:
    """Centre crop of the two in-plane axes (suppTools.py:399-422)."""
    size = data.shape[-1]
    start = (size - croped_size) // 2
    stop = croped_size + start
    return data[..., start:stop, start:stop]


def apply_circular_mask(data: torch.Tensor, recon_mask_radius: float, cupyrun: bool = True) -> torch.Tensor:
    """Zero everything outside the disc, in place (suppTools.py:364-396)."""
    if not data.is_contiguous():
        raise ValueError("apply_circular_mask needs a contiguous volume")
    n = data.shape[-1]
    nz = data.shape[0] if data.ndim == 3 else 1
    with torch.cuda.device(data.device):
        check(lib.tmb_circular_mask(ptr(data), nz, n, float(recon_mask_radius), stream_ptr(data)),
              "tmb_circular_mask")
    return data


def check_kwargs(reconstruction: torch.Tensor, **kwargs) -> torch.Tensor:
    """suppTools.py:462-467."""
    for key, value in kwargs.items():
        if key == "recon_mask_radius" and value is not None:
            apply_circular_mask(reconstruction, value, kwargs.get("cupyrun", True))
    return reconstruction
