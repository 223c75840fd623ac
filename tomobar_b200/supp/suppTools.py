"""Detector padding, reconstruction cropping, circular masking and flat/dark-field normalisation on
CUDA tensors (behaviour of tomobar/supp/suppTools.py:187-264, 364-467)."""

from __future__ import annotations

import numpy as np
import torch

from tomobar_b200._lib import lib, check
from tomobar_b200._tensors import ptr, stream_ptr


def _apply_horiz_detector_padding(data: torch.Tensor, detector_width_pad: int, cupyrun: bool = True) -> torch.Tensor:
    """Edge-pad detX (the last axis) on both sides (suppTools.py:425-459)."""
    if detector_width_pad <= 0:
        return data
    return edge_pad(data, int(detector_width_pad), data.shape[-1] + 2 * int(detector_width_pad))


def edge_pad(data: torch.Tensor, pad_left: int, width_out: int) -> torch.Tensor:
    """out[..., j] = data[..., clamp(j - pad_left, 0, w - 1)] in one fused pass (k_edge_pad)."""
    data = data.contiguous()
    w = data.shape[-1]
    out = torch.empty(tuple(data.shape[:-1]) + (int(width_out),), dtype=torch.float32, device=data.device)
    rows = data.numel() // w
    with torch.cuda.device(data.device):
        check(lib.tmb_edge_pad(ptr(data), ptr(out), rows, w, int(width_out), int(pad_left), stream_ptr(data)),
              "tmb_edge_pad")
    return out


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


def normaliser(data, flats, darks, log: bool = True, method: str = "mean", axis: int = 0, device=0,
               **kwargs) -> torch.Tensor:
    """Flat / dark-field normalisation with negative log (suppTools.py:187-264) in one fused pass
    from the raw (uint16 or float32) projections to the float32 CUDA sinogram.

    ``data`` is 3-D with the angle axis ``axis`` (0 or 1); ``flats`` / ``darks`` are stacks along the
    same axis (``darks`` may be None).  Methods "mean" and "median"; the reference's "dynamic" flat
    fielding (a CPU pre-processing routine) is not part of the hot path.
    """
    dev = torch.device("cuda", device) if not isinstance(device, torch.device) else device

    def to_dev(a):
        if isinstance(a, np.ndarray):
            a = torch.from_numpy(np.ascontiguousarray(a))
        return a.to(dev)

    data = to_dev(data)
    if data.ndim == 2:
        raise NameError("Normalisation is implemented for 3d data input")
    if axis not in (0, 1):
        raise ValueError("normaliser: the angle axis must be 0 or 1")
    flats = to_dev(flats).to(torch.float32)
    darks = torch.zeros_like(flats) if darks is None else to_dev(darks).to(torch.float32)
    if method is None or method == "mean":
        flat_m, dark_m = flats.mean(axis), darks.mean(axis)
    elif method == "median":
        flat_m, dark_m = flats.quantile(0.5, dim=axis), darks.quantile(0.5, dim=axis)
    elif method == "dynamic":
        raise NotImplementedError("dynamic flat-field correction is outside the GPU hot path")
    else:
        raise NameError("Please select an appropriate method for normalisation: mean, median or dynamic")
    is_u16 = data.dtype == torch.uint16
    if not is_u16:
        data = data.to(torch.float32)
    data = data.contiguous()
    flat_m, dark_m = flat_m.contiguous(), dark_m.contiguous()
    n0, n1, n2 = data.shape
    out = torch.empty((n0, n1, n2), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.tmb_normalise(ptr(data), int(is_u16), ptr(flat_m), ptr(dark_m), ptr(out), n0, n1, n2, int(axis),
                                int(bool(log)), stream_ptr(out)), "tmb_normalise")
    return out
