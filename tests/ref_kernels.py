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


def ref_FOURIER_INV(data, angles, recon_size, cor=0.0, pad=0, filter_type="shepp", cutoff_freq=1.0,
                    use_naive_prune=False, center_size=32768):
    """Default path of RecToolsDIRCuPy.FOURIER_INV (methodsDIR_CuPy.py:152-447) with the
    reference's own fft_us_kernels.cu kernels and launch geometry; data [detY, angles, detX]."""
    import math
    from tomobar_b200.fourier import calc_filter  # host-side numpy filter (pinned by test_fourier_filters)

    mod = module("fft_us_kernels")
    dev = data.device
    nz, nproj, data_n = data.shape
    odd_h, odd_v = bool(data_n % 2), bool(nz % 2)
    data_n += odd_h
    nz += odd_v
    if odd_h or odd_v:
        dp = torch.zeros((nz, nproj, data_n), dtype=torch.float32, device=dev)
        dp[: nz - odd_v, :, : data_n - odd_h] = data
        dp[: nz - odd_v, :, -int(odd_h)] = data[..., -int(odd_h)]
        data = dp
    n = data_n + 2 * pad
    center_size = min(center_size, 2 * n)
    theta = torch.as_tensor(-np.asarray(angles), dtype=torch.float32, device=dev)
    sidx = torch.argsort(theta)
    sth = theta[sidx].contiguous()
    sth_cpu = sth.cpu().numpy()
    pi_count = 1 + int(np.ceil(abs(sth_cpu[nproj - 1] - sth_cpu[0]) / math.pi))
    angle_range = torch.zeros((max(center_size, 1), max(center_size, 1), 1 + pi_count * 2), dtype=torch.int16, device=dev)
    eps = 1e-4
    mu = -np.log(eps) / (2 * n * n)
    # filtering (:449-545)
    over = 2 ** math.ceil(math.log2(data_n * 3))
    if n > over:
        over = 2 ** math.ceil(math.log2(n))
    padding_m = over // 2 - data_n // 2
    unpad_m, unpad_p = over // 2 - n // 2, over // 2 + n // 2
    wf = torch.as_tensor(calc_filter(over, filter_type, cutoff_freq), device=dev)
    t = torch.fft.rfftfreq(over, device=dev).to(torch.float32)
    w = wf * torch.exp(-2 * np.pi * 1j * t * (cor + 0.5))
    tmp_p = torch.empty((nz, nproj, n), dtype=torch.float32, device=dev)
    for z in range(nz):
        tmp = torch.nn.functional.pad(data[z:z + 1], (padding_m, padding_m), mode="replicate")
        tmp = torch.fft.irfft(w * torch.fft.rfft(tmp, dim=2), dim=2)
        tmp_p[z] = tmp[0, :, unpad_m:unpad_p]
    nz2 = nz // 2
    datac = torch.empty((nz2, nproj, n), dtype=torch.complex64, device=dev)
    # (zeros: the scatter branches add into it, methodsDIR_CuPy.py:661-670)
    fde = torch.zeros((nz2, 2 * n, 2 * n), dtype=torch.complex64, device=dev)
    i32 = np.int32
    cdiv = lambda a, b: int(np.ceil(a / b))
    mod.launch("r2c_c1dfftshift", (cdiv(n, 32), cdiv(nproj, 32), nz2), (32, 32, 1), [tmp_p, datac, i32(n), i32(nproj), i32(nz2)])
    datac = torch.fft.fft(datac, dim=-1).contiguous()
    m = int(np.ceil(2 * n * 1 / np.pi * np.sqrt(-mu * np.log(eps) + (mu * n) * (mu * n) / 4)))
    mod.launch("c1dfftshift", (cdiv(n, 32), cdiv(nproj, 32), nz2), (32, 32, 1),
               [datac, np.float32(4 / n), i32(n), i32(nproj), i32(nz2)])
    if center_size >= 192:  # _CENTER_SIZE_MIN (methodsDIR_CuPy.py:23, 759-816)
        if center_size != 2 * n:
            mod.launch("gather_kernel_partial", (cdiv(n, 16), cdiv(nproj, 16), nz2), (16, 16, 1),
                       [datac, fde, theta, i32(m), np.float32(mu), i32(center_size), i32(n), i32(nproj), i32(nz2)])
        prune = "gather_kernel_center_prune_naive" if use_naive_prune else "gather_kernel_center_angle_based_prune"
        mod.launch(prune, (cdiv(center_size, 256), center_size, 1), (256, 1, 1),
                   [angle_range, i32(pi_count * 2 + 1), sth, i32(m), i32(center_size), i32(n), i32(nproj)])
        mod.launch("gather_kernel_center", (cdiv(center_size, 32), cdiv(center_size, 4), nz2), (32, 4, 1),
                   [datac, fde, angle_range, i32(pi_count * 2 + 1), theta, sidx.to(torch.int64).contiguous(), i32(m),
                    np.float32(mu), i32(center_size), i32(n), i32(nproj), i32(nz2)])
    else:  # :818-835
        mod.launch("gather_kernel", (cdiv(n, 16), cdiv(nproj, 16), nz2), (16, 16, 1),
                   [datac, fde, theta, i32(m), np.float32(mu), i32(n), i32(nproj), i32(nz2)])
    mod.launch("c2dfftshift", (cdiv(2 * n, 32), cdiv(2 * n, 8), nz2), (32, 8, 1), [fde, i32(n), i32(nz2)])
    for z in range(nz2):
        fde[z] = torch.fft.ifft2(fde[z])
    mod.launch("c2dfftshift", (cdiv(2 * n, 32), cdiv(2 * n, 8), nz2), (32, 8, 1), [fde, i32(n), i32(nz2)])
    odd_r = bool(recon_size % 2)
    unpad_z = nz - odd_v
    um = (n - odd_h) // 2 - recon_size // 2
    up = (n - odd_h) // 2 + (recon_size + odd_r) // 2
    rs = up - um
    recon = torch.empty((unpad_z, rs, rs), dtype=torch.float32, device=dev)
    mod.launch("unpadding_mul_phi", (cdiv(rs, 32), cdiv(rs, 32), nz2), (32, 32, 1),
               [recon, fde, np.float32(mu), i32(nproj), i32(up), i32(unpad_z), i32(um), i32(n), i32(nz2)])
    torch.cuda.synchronize()
    return recon
