"""Small helpers around torch tensors used as device-memory handles."""

from __future__ import annotations

import numpy as np
import torch


def as_cuda_f32(x, device=None, what: str = "array") -> torch.Tensor:
    """Accept a torch tensor, a numpy array or anything exposing ``__cuda_array_interface__`` /
    DLPack (e.g. a CuPy array) and return a float32 CUDA torch tensor (no copy when possible)."""
    if isinstance(x, torch.Tensor):
        t = x
    elif isinstance(x, np.ndarray):
        t = torch.from_numpy(x)
    elif hasattr(x, "__dlpack__"):
        t = torch.from_dlpack(x)
    elif hasattr(x, "__cuda_array_interface__"):
        t = torch.as_tensor(x, device="cuda")
    else:
        t = torch.as_tensor(np.asarray(x))
    if t.dtype != torch.float32:
        raise ValueError(f"The {what} should be float32 data type")
    if not t.is_cuda:
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        t = t.to(device)
    return t


def ptr(t: torch.Tensor) -> int:
    return t.data_ptr()


def stream_ptr(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def require_dense(t: torch.Tensor, shape, what: str) -> torch.Tensor:
    if tuple(t.shape) != tuple(shape):
        raise ValueError(f"{what} has shape {tuple(t.shape)}, expected {tuple(shape)}")
    if not t.is_contiguous():
        # the reference hands ASTRA the raw pointer of a possibly strided view
        # (astra_base.py:533-535); here the logical array is used
        t = t.contiguous()
    return t
