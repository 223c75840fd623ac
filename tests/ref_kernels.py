"""TEST INFRASTRUCTURE: runs the reference's OWN raw CUDA kernels (compiled by
oracle/build_ref.sh from /root/reference/tomobar/cuda_kernels/*.cu into oracle/_ref/*.cubin)
on the GPU box through the CUDA driver API, with the launch geometry and ping-pong loops of the
reference's host code restated here (regularisersCuPy.py:41-296, fourier.py:52-66).

This is the "real reference" for kernel-level parity of libtmb's TV / filter kernels; it is
never imported by the product package.
"""

from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "primal_dual_for_total_variation.cubin"))


class RefModule:
    def __init__(self, name: str):
        from cuda.bindings import driver as cu

        self.cu = cu
        torch.zeros(1, device="cuda")  # make torch's primary context current
        with open(os.path.join(REF_DIR, name + ".cubin"), "rb") as f:
            image = f.read()
        err, self.mod = cu.cuModuleLoadData(image)
        assert err == cu.CUresult.CUDA_SUCCESS, err
        self._fn = {}

    def launch(self, fname, grid, block, args, shared_mem=0):
        cu = self.cu
        if fname not in self._fn:
            err, f = cu.cuModuleGetFunction(self.mod, fname.encode())
            assert err == cu.CUresult.CUDA_SUCCESS, (fname, err)
            self._fn[fname] = f
        vals, types = [], []
        for a in args:
            if isinstance(a, torch.Tensor):
                vals.append(a.data_ptr()); types.append(ctypes.c_void_p)
            elif isinstance(a, (np.float32, float)):
                vals.append(float(a)); types.append(ctypes.c_float)
            elif isinstance(a, (np.integer, int)):
                vals.append(int(a)); types.append(ctypes.c_int)
            else:
                raise TypeError(type(a))
        grid = tuple(grid) + (1,) * (3 - len(grid))
        block = tuple(block) + (1,) * (3 - len(block))
        stream = torch.cuda.current_stream().cuda_stream
        (err,) = cu.cuLaunchKernel(self._fn[fname], *[int(g) for g in grid], *[int(b) for b in block],
                                   int(shared_mem), stream, (tuple(vals), tuple(types)), 0)
        assert err == cu.CUresult.CUDA_SUCCESS, (fname, err)


_MODS = {}


def module(name):
    if name not in _MODS:
        _MODS[name] = RefModule(name)
    return _MODS[name]


def _squeeze(data):
    if data.ndim == 2:
        return data, True, 0
    for i in range(3):
        if data.shape[i] == 1:
            return data.squeeze(i), True, i
    return data, False, 0


def ref_PD_TV(data, regularisation_parameter=1e-5, iterations=1000, methodTV=0, nonneg=0, lipschitz_const=8.0,
              half_precision=False):
    """Host loop of PD_TV_cupy (regularisersCuPy.py:170-296) around the reference's own kernels."""
    data, is2d, ax = _squeeze(data)
    data = data.contiguous()
    pdt = torch.float16 if half_precision else torch.float32
    tau = np.float32(regularisation_parameter * 0.1)
    sigma = np.float32(1.0 / (lipschitz_const * tau))
    theta = np.float32(1.0)
    lt = np.float32(tau / regularisation_parameter)
    U = [data.clone(), torch.zeros_like(data)]
    nd = data.ndim
    P = [[torch.zeros(data.shape, dtype=pdt, device=data.device) for _ in range(2)] for _ in range(nd)]
    name = f"primal_dual_for_total_variation_{'3D' if nd == 3 else '2D'}_{'half' if half_precision else 'float'}"
    if nonneg:
        name += "_nonneg"
    if methodTV:
        name += "_methodTV"
    mod = module("primal_dual_for_total_variation")
    dz, dy, dx = (0,) * (3 - nd) + tuple(data.shape)
    grid = ((dx + 127) // 128, dy) + ((dz,) if nd == 3 else ())
    dims = (dx, dy) + ((dz,) if nd == 3 else ())
    i, o = 0, 1
    for _ in range(iterations):
        args = [data, U[i], U[o]] + [P[d][i] for d in range(nd)] + [P[d][o] for d in range(nd)] + \
               [sigma, tau, lt, theta] + [np.int32(v) for v in dims]
        mod.launch(name, grid, (128, 1, 1), args)
        i, o = o, i
    out = U[i]
    return out.unsqueeze(ax) if is2d else out


def ref_ROF_TV(data, regularisation_parameter=1e-5, iterations=3000, time_marching_parameter=0.001,
               half_precision=False):
    """Host loop of ROF_TV_cupy (regularisersCuPy.py:41-167) around the reference's own kernels."""
    data, is2d, ax = _squeeze(data)
    data = data.contiguous()
    ddt = torch.float16 if half_precision else torch.float32
    nd = data.ndim
    U = [data.clone(), torch.zeros_like(data)]
    D = [torch.empty(data.shape, dtype=ddt, device=data.device) for _ in range(nd)]
    mod = module("rudin_osher_fatemi_total_variation")
    suffix = f"{nd}D_{'half' if half_precision else 'float'}"
    dz, dy, dx = (0,) * (3 - nd) + tuple(data.shape)
    grid = ((dx + 127) // 128, dy) + ((dz,) if nd == 3 else ())
    dims = [np.int32(v) for v in ((dx, dy) + ((dz,) if nd == 3 else ()))]
    i, o = 0, 1
    for _ in range(iterations):
        mod.launch("divergence_kernel_" + suffix, grid, (128, 1, 1), [U[i]] + D + dims)
        mod.launch("TV_kernel_" + suffix, grid, (128, 1, 1),
                   [U[i], U[o], data] + D + [np.float32(regularisation_parameter),
                                             np.float32(time_marching_parameter)] + dims)
        i, o = o, i
    out = U[i]
    return out.unsqueeze(ax) if is2d else out


def ref_filtersinc(n, cutoff, multiplier, device="cuda"):
    """generate_filtersinc launch of fourier.py:52-66."""
    f = torch.empty(n // 2 + 1, dtype=torch.float32, device=device)
    module("generate_filtersync").launch("generate_filtersinc", (1, 1, 1), (256, 1, 1),
                                         [np.float32(cutoff), f, np.int32(n), np.float32(multiplier)],
                                         shared_mem=256 * 4)
    return f
