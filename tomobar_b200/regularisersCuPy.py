"""TV proximal operators on sm_100a (libtmb.so) with the reference's call signatures.

``ROF_TV_cupy`` / ``PD_TV_cupy`` / ``prox_regul`` keep the names, argument order, defaults
and error behaviour of tomobar/regularisersCuPy.py:6-315; arrays are float32 CUDA torch
tensors (CuPy arrays are accepted through DLPack).
"""

from __future__ import annotations

from typing import Tuple

import torch

from tomobar_b200._lib import lib, check
from tomobar_b200._tensors import as_cuda_f32, ptr, stream_ptr


def prox_regul(self, X, _regularisation_: dict) -> torch.Tensor:
    """regularisersCuPy.py:6-38 -- `self` is the reconstruction object (reads
    ``self.Atools.device_index`` and ``self.nonneg_regul``)."""
    method = _regularisation_["method"]
    if "ROF_TV" in method:
        return ROF_TV_cupy(
            X,
            _regularisation_["regul_param"],
            _regularisation_["iterations"],
            _regularisation_["time_marching_step"],
            self.Atools.device_index,
            _regularisation_.get("half_precision", False),
        )
    if "PD_TV" in method:
        return PD_TV_cupy(
            X,
            _regularisation_["regul_param"],
            _regularisation_["iterations"],
            _regularisation_["methodTV"],
            self.nonneg_regul,
            _regularisation_["PD_LipschitzConstant"],
            self.Atools.device_index,
            _regularisation_.get("half_precision", False),
        )
    raise ValueError(f"Unknown regularisation method {method!r}: ROF_TV and PD_TV are supported")


def _squeeze_unit_axis(data: torch.Tensor) -> Tuple[torch.Tensor, bool, int]:
    """2-D input, or 3-D input with a unit axis, runs through the 2-D kernels
    (regularisersCuPy.py:299-315)."""
    if data.ndim == 2:
        return data, True, 0
    if data.ndim == 3:
        for axis in range(3):
            if data.shape[axis] == 1:
                return data.squeeze(axis), True, axis
        return data, False, 0
    raise ValueError("2D or 3D arrays must be provided only")


def _prepare(data, gpu_id: int):
    if gpu_id < 0:
        raise ValueError("The gpu_device must be a positive integer or zero")
    if isinstance(data, torch.Tensor) and data.dtype != torch.float32:
        raise ValueError("The input data should be float32 data type")
    data = as_cuda_f32(data, torch.device("cuda", gpu_id), "input data")
    data, is2d, axis = _squeeze_unit_axis(data)
    data = data.contiguous()
    dz, dy, dx = (1,) * (3 - data.ndim) + tuple(data.shape)
    return data, is2d, axis, dz, dy, dx


def _result_buffer(data: torch.Tensor, is2d: bool, axis: int, out):
    """Kernel output buffer (shape of the squeezed data) and the tensor handed back to the caller;
    2-D results come back re-expanded on the squeezed axis (regularisersCuPy.py:164-167)."""
    if out is None:
        res = torch.empty_like(data)
        return res, (res.unsqueeze(axis) if is2d else res)
    expected = tuple(data.unsqueeze(axis).shape) if is2d else tuple(data.shape)
    if tuple(out.shape) != expected or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous float32 tensor of shape {expected}")
    return (out.squeeze(axis) if is2d else out), out


def ROF_TV_cupy(
    data,
    regularisation_parameter: float = 1e-05,
    iterations: int = 3000,
    time_marching_parameter: float = 0.001,
    gpu_id: int = 0,
    half_precision: bool = False,
    out: torch.Tensor = None,
) -> torch.Tensor:
    """regularisersCuPy.py:41-167."""
    data, is2d, axis, dz, dy, dx = _prepare(data, gpu_id)
    res, ret = _result_buffer(data, is2d, axis, out)
    ws = torch.empty(lib.tmb_tv_workspace_bytes(1, dz, dy, dx, int(half_precision)), dtype=torch.uint8,
                     device=data.device)
    with torch.cuda.device(data.device):
        check(lib.tmb_rof_tv(ptr(data), ptr(res), dz, dy, dx, float(regularisation_parameter), int(iterations),
                             float(time_marching_parameter), int(bool(half_precision)), ptr(ws), stream_ptr(data)),
              "tmb_rof_tv")
    return ret


def PD_TV_cupy(
    data,
    regularisation_parameter: float = 1e-05,
    iterations: int = 1000,
    methodTV: int = 0,
    nonneg: int = 0,
    lipschitz_const: float = 8.0,
    gpu_id: int = 0,
    half_precision: bool = False,
    out: torch.Tensor = None,
) -> torch.Tensor:
    """regularisersCuPy.py:170-296."""
    data, is2d, axis, dz, dy, dx = _prepare(data, gpu_id)
    res, ret = _result_buffer(data, is2d, axis, out)
    ws = torch.empty(lib.tmb_tv_workspace_bytes(0, dz, dy, dx, int(half_precision)), dtype=torch.uint8,
                     device=data.device)
    with torch.cuda.device(data.device):
        check(lib.tmb_pd_tv(ptr(data), ptr(res), dz, dy, dx, float(regularisation_parameter), int(iterations),
                            int(methodTV), int(nonneg), float(lipschitz_const), int(bool(half_precision)), ptr(ws),
                            stream_ptr(data)), "tmb_pd_tv")
    return ret
