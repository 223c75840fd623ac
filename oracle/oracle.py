"""ORACLE -- TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy + the small C file next to it) of the reference hot path.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product package
``tomobar_b200`` never does (tests/test_no_oracle_in_product.py enforces it).

Every function cites the reference ``file:line`` it follows (paths relative to the
reference checkout).  The projector arithmetic itself lives in astra-toolbox==2.4.*
(pyproject.toml:41, not vendored); its published par3d model is restated in
``proj_oracle.c`` per SURVEY.md Appendix A and pinned against the reference's own test
goldens (tests/test_oracle_goldens.py).
"""

from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_LIB = None

f32 = np.float32


_SOURCES = ("proj_oracle.c", "tv_oracle.c", "fbp2d_oracle.c")


def build(force: bool = False) -> str:
    """Compile the C restatements -> oracle/_build/liboracle.so (gcc -O2 -fopenmp)."""
    so = os.path.join(_BUILD, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in _SOURCES]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs):
        os.makedirs(_BUILD, exist_ok=True)
        # -ffp-contract=off: every fused multiply-add in the oracle is an explicit fmaf()
        subprocess.check_call(
            ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
             "-o", so] + srcs + ["-lm"]
        )
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        fp = ctypes.POINTER(ctypes.c_float)
        for name in ("oracle_bp3d", "oracle_fp3d"):
            fn = getattr(_LIB, name)
            fn.restype = None
            fn.argtypes = [fp, fp, fp] + [ctypes.c_int] * 5
        ci, cf = ctypes.c_int, ctypes.c_float
        _LIB.oracle_pd_tv.restype = ci
        _LIB.oracle_pd_tv.argtypes = [fp, fp, ci, ci, ci, cf, cf, cf, cf, ci, ci, ci]
        _LIB.oracle_rof_tv.restype = ci
        _LIB.oracle_rof_tv.argtypes = [fp, fp, ci, ci, ci, cf, ci, cf]
        _LIB.oracle_bp2d_line.restype = None
        _LIB.oracle_bp2d_line.argtypes = [fp, fp, ctypes.POINTER(ctypes.c_double), ci, ci, ci, ci]
        _LIB.oracle_threads.restype = ci
        _LIB.oracle_threads.argtypes = []
        _LIB.oracle_set_threads.restype = None
        _LIB.oracle_set_threads.argtypes = [ci]
    return _LIB


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


# --------------------------------------------------------------------------------------
# geometry  (supp/funcs.py:45-81, astra_base.py:195-209, 215-222, 244-255, 287-308)
# --------------------------------------------------------------------------------------
def angle_table(angles: np.ndarray, cor, n: int, nu: int) -> np.ndarray:
    """Per-angle fp32 table [na, 8] (see proj_oracle.c header).

    sin/cos are evaluated in the dtype of ``angles`` exactly like
    supp/funcs.py:74-81 (``np.cos(theta)`` on a numpy scalar), everything else is derived
    in double and rounded once to fp32 (ASTRA derives its kernel constants in double from
    the ``parallel3d_vec`` vectors).
    """
    angles = np.asarray(angles)
    na = angles.size
    ca = np.cos(angles).astype(np.float64)
    sa = np.sin(angles).astype(np.float64)
    cor = np.broadcast_to(np.asarray(cor, dtype=np.float64), (na,))
    tbl = np.zeros((na, 8), dtype=np.float64)
    tbl[:, 0] = ca
    tbl[:, 1] = sa
    tbl[:, 2] = -cor + (nu / 2.0 - 0.5)
    dirx = np.abs(sa) > np.abs(ca)  # march along x (columns), interpolate along rows
    major = np.where(dirx, sa, ca)
    minor = np.where(dirx, ca, sa)
    alpha = -minor / major
    tbl[:, 3] = alpha
    tbl[:, 4] = (-nu / 2.0 + 0.5 + cor) / major + (n / 2.0 - 0.5)
    tbl[:, 5] = 1.0 / major
    tbl[:, 6] = np.sqrt(1.0 + alpha * alpha)
    tbl[:, 7] = np.where(dirx, 0.0, 1.0)
    return np.ascontiguousarray(tbl.astype(np.float32))


def os_indices(na: int, os_number: int):
    """astra_base.py:195-209 -- zero-padded interleaved subset table + bins."""
    bins = int(np.ceil(float(na) / float(os_number)))
    tab = np.zeros([os_number, bins], dtype="int")
    for s in range(os_number):
        sel = 0
        for p in range(bins):
            idx = sel + s
            if idx < na:
                tab[s, p] = idx
                sel += os_number
    return tab, bins


def subset_indices(tab: np.ndarray, bins: int, s: int) -> np.ndarray:
    """Consumers drop ONE trailing entry when it is 0 (methodsIR_CuPy.py:454-456)."""
    ind = tab[s, :]
    if ind[bins - 1] == 0:
        ind = ind[:-1]
    return ind


def fp3d(vol: np.ndarray, tbl: np.ndarray, nu: int, quant: bool = True) -> np.ndarray:
    vol = np.ascontiguousarray(vol, dtype=np.float32)
    nz, n, _ = vol.shape
    na = tbl.shape[0]
    tbl = np.ascontiguousarray(tbl, dtype=np.float32)
    out = np.empty((nz, na, nu), dtype=np.float32)
    _lib().oracle_fp3d(_ptr(vol), _ptr(out), _ptr(tbl), nz, n, nu, na, int(quant))
    return out


def bp3d(sino: np.ndarray, tbl: np.ndarray, n: int, quant: bool = True) -> np.ndarray:
    sino = np.ascontiguousarray(sino, dtype=np.float32)
    nz, na, nu = sino.shape
    assert na == tbl.shape[0]
    tbl = np.ascontiguousarray(tbl, dtype=np.float32)
    out = np.empty((nz, n, n), dtype=np.float32)
    _lib().oracle_bp3d(_ptr(sino), _ptr(out), _ptr(tbl), nz, n, nu, na, int(quant))
    return out


class Atools:
    """Oracle stand-in for ``AstraTools3D`` (astra_tools3d.py:19-110)."""

    def __init__(self, detectors_x, detectors_x_pad, detectors_y, angles_vec, centre_of_rotation,
                 recon_size, ordsub_number=None, quant=True):
        self.detectors_x = detectors_x
        self.detectors_x_pad = detectors_x_pad
        self.detectors_y = detectors_y
        self.angles_vec = np.asarray(angles_vec)
        self.centre_of_rotation = 0.0 if centre_of_rotation is None else centre_of_rotation
        self.recon_size = recon_size
        self.ordsub_number = 1 if ordsub_number is None else ordsub_number
        self.quant = quant
        self.nu = detectors_x + 2 * detectors_x_pad
        self.vol_shape = (detectors_y, recon_size, recon_size)
        self.tbl = angle_table(self.angles_vec, self.centre_of_rotation, recon_size, self.nu)
        if self.ordsub_number > 1:
            self.newInd_Vec, self.NumbProjBins = os_indices(self.angles_vec.size, self.ordsub_number)
            self.tbl_os = [self.tbl[subset_indices(self.newInd_Vec, self.NumbProjBins, s)]
                           for s in range(self.ordsub_number)]

    def _forwprojCuPy(self, x):
        return fp3d(x, self.tbl, self.nu, self.quant)

    def _backprojCuPy(self, b):
        return bp3d(b, self.tbl, self.recon_size, self.quant)

    def _forwprojOSCuPy(self, x, os_index):
        return fp3d(x, self.tbl_os[os_index], self.nu, self.quant)

    def _backprojOSCuPy(self, b, os_index):
        return bp3d(b, self.tbl_os[os_index], self.recon_size, self.quant)


# --------------------------------------------------------------------------------------
# helpers  (supp/suppTools.py:364-459)
# --------------------------------------------------------------------------------------
def pad_detector(data: np.ndarray, pad: int) -> np.ndarray:
    """suppTools.py:425-459 (edge padding of detX, last axis)."""
    if pad <= 0:
        return data
    width = [(0, 0)] * (data.ndim - 1) + [(pad, pad)]
    return np.pad(data, width, mode="edge")


def circular_mask(data: np.ndarray, radius: float) -> np.ndarray:
    """suppTools.py:364-396 (in place, like the reference)."""
    n = data.shape[-1]
    h = n // 2
    Y, X = np.ogrid[:n, :n]
    dist = np.sqrt((X - h) ** 2 + (Y - h) ** 2)
    if radius <= 1.0:
        mask = dist <= h - abs(h - h / radius)
    else:
        mask = dist <= h + abs(h - h / radius)
    data *= mask
    return data


def recon_crop(data: np.ndarray, size: int) -> np.ndarray:
    """suppTools.py:399-422."""
    n = data.shape[-1]
    s = (n - size) // 2
    return data[..., s:s + size, s:s + size]


# --------------------------------------------------------------------------------------
# TV proximal operators (literal restatements of the two .cu files)
# --------------------------------------------------------------------------------------
def _squeeze_2d(data):
    """regularisersCuPy.py:299-315."""
    if data.ndim == 2:
        return data, True, 0
    for i in range(3):
        if data.shape[i] == 1:
            return np.squeeze(data, axis=i), True, i
    return data, False, 0


def _fwd(U, axis):
    """forward difference, ``U[-1]-U`` at the last index
    (primal_dual_for_total_variation.cu:214-222)."""
    nxt = np.roll(U, -1, axis=axis)
    sl = [slice(None)] * U.ndim
    sl[axis] = -1
    if U.shape[axis] > 1:
        sl2 = list(sl)
        sl2[axis] = -2
        nxt[tuple(sl)] = U[tuple(sl2)]
    else:
        nxt[tuple(sl)] = 0.0
    return nxt - U


def _bwd0(P, axis):
    """P - P[-1] with P[-1] := 0 at index 0 (primal_dual_for_total_variation.cu:147-162)."""
    prv = np.roll(P, 1, axis=axis)
    sl = [slice(None)] * P.ndim
    sl[axis] = 0
    prv[tuple(sl)] = 0.0
    return P - prv


def threads() -> int:
    """OpenMP threads the C restatements run on (what ``cores`` of a CPU baseline reports)."""
    return int(_lib().oracle_threads())


def set_threads(n: int) -> int:
    """Size the OpenMP pool explicitly (an inherited OMP_NUM_THREADS=1, e.g. from torchrun, must not decide it)."""
    _lib().oracle_set_threads(int(n))
    return threads()


def pd_tv(data, regularisation_parameter=1e-5, iterations=1000, methodTV=0, nonneg=0,
          lipschitz_const=8.0, half_precision=False, use_c=True):
    """regularisersCuPy.py:170-296 + primal_dual_for_total_variation.cu:126-261 / :361-492.

    ``use_c`` (fp32 duals): the OpenMP twin in tv_oracle.c, bit-identical to the numpy statements below
    (tests/test_oracle_tv_c.py) -- the numpy version is single-threaded."""
    data = np.asarray(data)
    if data.dtype != np.float32:
        raise ValueError("The input data should be float32 data type")
    data, is2d, ax = _squeeze_2d(data)
    tau = f32(regularisation_parameter * 0.1)
    sigma = f32(1.0 / (lipschitz_const * tau))
    theta = f32(1.0)
    lt = f32(tau / regularisation_parameter)
    nd = data.ndim
    if use_c and not half_precision:
        src = np.ascontiguousarray(data)
        shp = (1,) * (3 - nd) + src.shape
        out = np.empty_like(src)
        if _lib().oracle_pd_tv(_ptr(src), _ptr(out), shp[0], shp[1], shp[2], float(tau), float(sigma), float(lt),
                               float(theta), int(iterations), int(methodTV), int(nonneg)) != 0:
            raise MemoryError("oracle_pd_tv")
        return np.expand_dims(out, ax) if is2d else out
    U = data.copy()
    pdt = np.float16 if half_precision else np.float32
    P = [np.zeros(data.shape, dtype=pdt) for _ in range(nd)]
    axes = [nd - 1 - d for d in range(nd)]  # P1 <-> fast axis, P2 middle, P3 slow
    for _ in range(iterations):
        Pn = [P[d].astype(np.float32) + sigma * _fwd(U, axes[d]) for d in range(nd)]
        if methodTV == 0:
            den = sum(p * p for p in Pn)
            with np.errstate(divide="ignore"):
                sc = np.where(den > 1.0, f32(1.0) / np.sqrt(den), f32(1.0)).astype(np.float32)
            Pn = [p * sc for p in Pn]
        else:
            Pn = [p / np.maximum(np.abs(p), f32(1.0)) for p in Pn]
        div = sum(-_bwd0(Pn[d], axes[d]) for d in range(nd))
        Ub = np.maximum(U, f32(0.0)) if nonneg else U
        newU = (Ub - tau * div + lt * data) / (f32(1.0) + lt)
        U = (newU + theta * (newU - Ub)).astype(np.float32)
        P = [p.astype(pdt) for p in Pn]
    return np.expand_dims(U, ax) if is2d else U


def _refl_next(U, axis):
    idx = np.arange(U.shape[axis]) + 1
    idx[-1] = U.shape[axis] - 2 if U.shape[axis] > 1 else 0
    return np.take(U, idx, axis=axis)


def _refl_prev(U, axis):
    idx = np.arange(U.shape[axis]) - 1
    idx[0] = 1 if U.shape[axis] > 1 else 0
    return np.take(U, idx, axis=axis)


def rof_tv(data, regularisation_parameter=1e-5, iterations=3000, time_marching_parameter=0.001,
           half_precision=False, use_c=True):
    """regularisersCuPy.py:41-167 + rudin_osher_fatemi_total_variation.cu:70-148, 157-248.
    ``use_c``: as for ``pd_tv``."""
    data = np.asarray(data)
    if data.dtype != np.float32:
        raise ValueError("The input data should be float32 data type")
    data, is2d, ax = _squeeze_2d(data)
    nd = data.ndim
    lam = f32(regularisation_parameter)
    tau = f32(time_marching_parameter)
    if use_c and not half_precision:
        src = np.ascontiguousarray(data)
        shp = (1,) * (3 - nd) + src.shape
        out = np.empty_like(src)
        if _lib().oracle_rof_tv(_ptr(src), _ptr(out), shp[0], shp[1], shp[2], float(lam), int(iterations),
                                float(tau)) != 0:
            raise MemoryError("oracle_rof_tv")
        return np.expand_dims(out, ax) if is2d else out
    ddt = np.float16 if half_precision else np.float32
    U = data.copy()
    for _ in range(iterations):
        nplus, m = [], []
        for axis in range(nd):
            n1 = _refl_next(U, axis) - U
            n0 = U - _refl_prev(U, axis)
            # calculate_denominator (:51-55): 0.5 is a double literal, result stored as float
            den = (0.5 * (np.sign(n1) + np.sign(n0)).astype(np.float64)
                   * np.minimum(np.abs(n1), np.abs(n0)).astype(np.float64)).astype(np.float32)
            nplus.append(n1)
            m.append(den * den)
        D = []
        for axis in range(nd):
            # normalize_difference (:57-61): float sums in argument order, + EPS in double
            terms = [nplus[e] * nplus[e] if e == axis else m[e] for e in range(nd)]
            # argument order in the kernels is (x=middle, y=fast, z=slow); keep it
            order = ([nd - 2, nd - 1] if nd == 2 else [1, 2, 0])
            s = terms[order[0]]
            for e in order[1:]:
                s = s + terms[e]
            s = (s.astype(np.float64) + 1.0e-8).astype(np.float32)
            D.append((nplus[axis] / np.sqrt(s)).astype(ddt).astype(np.float32))
        dv = None
        order = ([nd - 2, nd - 1] if nd == 2 else [1, 2, 0])
        for e in order:
            t = D[e] - _refl_prev(D[e], e)
            dv = t if dv is None else dv + t
        U = (U + tau * (lam * dv - (U - data))).astype(np.float32)
    return np.expand_dims(U, ax) if is2d else U


def prox_regul(X, reg: dict, nonneg_regul: int):
    """regularisersCuPy.py:6-38."""
    if "ROF_TV" in reg["method"]:
        return rof_tv(X, reg["regul_param"], reg["iterations"], reg["time_marching_step"],
                      reg.get("half_precision", False))
    if "PD_TV" in reg["method"]:
        return pd_tv(X, reg["regul_param"], reg["iterations"], reg["methodTV"], nonneg_regul,
                     reg["PD_LipschitzConstant"], reg.get("half_precision", False))
    raise ValueError("unknown regulariser")


def _reg_defaults(reg: Optional[dict]) -> dict:
    """supp/dicts.py:157-183."""
    reg = dict(reg or {})
    if not reg:
        reg["method"] = None
    reg.setdefault("regul_param", 0.001)
    reg.setdefault("iterations", 150)
    reg.setdefault("time_marching_step", 0.005)
    reg.setdefault("PD_LipschitzConstant", 12.0)
    reg.setdefault("methodTV", 0)
    return reg


# --------------------------------------------------------------------------------------
# FBP sinc filter (fourier.py:26-78, generate_filtersync.cu:5-82)
# --------------------------------------------------------------------------------------
def sinc_filter(n: int, cutoff: float, multiplier: float) -> np.ndarray:
    """Half-spectrum filter f[n//2+1] in fp32 as the one-block kernel builds it."""
    a = f32(cutoff)
    pi = f32(3.1415926535897932384626433832795)
    dw = f32(2) * pi / f32(n)
    i = np.arange(n, dtype=np.float32)
    w = (-pi + i * dw).astype(np.float32)
    rd = (a * w / f32(2.0)).astype(np.float32)
    sum_sq = f32(np.sum((rd * rd).astype(np.float64)))
    rn2 = np.sin(rd.astype(np.float64)).astype(np.float32)
    dot = f32(np.sum(((rn2 * rd) / sum_sq).astype(np.float64)))
    dot_sq = f32(dot * dot)
    rn1 = np.abs(2.0 / np.float64(a) * rn2.astype(np.float64)).astype(np.float32)
    r = (rn1 * dot_sq * f32(multiplier)).astype(np.float32)
    out = np.zeros(n // 2 + 1, dtype=np.float32)
    idx = (np.arange(n) + n // 2) % n
    keep = idx < n // 2 + 1
    out[idx[keep]] = r[keep]
    return out


def filtersinc3d(proj: np.ndarray, cutoff: float) -> np.ndarray:
    """fourier.py:26-78: proj[angles, detY, detX] -> irfft(rfft(p) * f)."""
    na, _, nu = proj.shape
    f = sinc_filter(nu, cutoff, 1.0 / na / nu)
    pf = np.fft.rfft(proj.astype(np.float32), axis=-1).astype(np.complex64)
    pf *= f
    # irfft(norm="forward") applies no scaling on the inverse
    return (np.fft.irfft(pf, nu, axis=-1) * nu).astype(np.float32)


# --------------------------------------------------------------------------------------
# the reference's CPU direct method for 2-D data (BASELINE.json config 1): methodsDIR.py:121-175, 295-320
# --------------------------------------------------------------------------------------
def filtersinc2d(sinogram: np.ndarray) -> np.ndarray:
    """methodsDIR.py:295-320: the numpy sinc filter of the CPU class (cut-off a = 1.1 hard-coded, full
    complex FFT per projection, scaled by 1 / number of projections)."""
    a = 1.1
    na, nu = sinogram.shape
    w = np.linspace(-np.pi, np.pi - (2 * np.pi) / nu, nu, dtype="float32")
    half = a * w / 2.0
    ramp = np.abs(2.0 / a * np.sin(half))
    # np.dot(rn2, pinv(rd as a 1 x n matrix)) == sum(sin(rd) * rd) / sum(rd^2)
    kappa = np.dot(np.sin(half), np.linalg.pinv(half.astype(np.float64)[None, :]))
    f = np.fft.fftshift(ramp * kappa ** 2)
    out = np.zeros(sinogram.shape)
    for i in range(na):
        out[i, :] = (1.0 / na) * np.real(np.fft.ifft(np.fft.fft(sinogram[i, :]) * f))
    return np.float32(out)


def parallel2d_vectors(angles: np.ndarray) -> np.ndarray:
    """[na, 6] (ray, detector centre, u) of ASTRA's classic "parallel" 2-D geometry, which is what the CPU
    class builds (astra_base.py:224-232: create_proj_geom("parallel", 1.0, detectors, angles); the centre of
    rotation is not part of it).  Same convention as supp/funcs.py:22-43 with a zero offset."""
    t = np.asarray(angles, dtype=np.float32).astype(np.float64)
    v = np.zeros((t.size, 6))
    v[:, 0], v[:, 1] = np.sin(t), -np.cos(t)
    v[:, 4], v[:, 5] = np.cos(t), np.sin(t)
    return v


def bp2d_line(sinogram: np.ndarray, angles: np.ndarray, n: int, threads: int = 1) -> np.ndarray:
    """ASTRA's CPU ``BP`` with the ``line`` projector (fbp2d_oracle.c); image row 0 is the TOP row."""
    sinogram = np.ascontiguousarray(sinogram, dtype=np.float32)
    na, nu = sinogram.shape
    vec = np.ascontiguousarray(parallel2d_vectors(angles))
    out = np.empty((n, n), np.float32)
    _lib().oracle_bp2d_line(_ptr(sinogram), _ptr(out), vec.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                            int(n), int(nu), int(na), int(threads))
    return out


def fbp2d_cpu(sinogram: np.ndarray, angles: np.ndarray, n: int, pad: int = 0, threads: int = 1) -> np.ndarray:
    """``RecToolsDIR(device_projector="cpu").FBP`` for 2-D data [angles, detX] (methodsDIR.py:161-168)."""
    data = pad_detector(np.asarray(sinogram, dtype=np.float32), pad)
    return bp2d_line(filtersinc2d(data), angles, n, threads)


# --------------------------------------------------------------------------------------
# reconstruction loops (methodsIR_CuPy.py, data_fidelities.py, methodsDIR_CuPy.py)
# --------------------------------------------------------------------------------------
class RecIR:
    """Oracle of ``RecToolsIRCuPy`` (methodsIR_CuPy.py:36-667). Data layout [detY, angles, detX]."""

    def __init__(self, DetectorsDimH, DetectorsDimH_pad, DetectorsDimV, CenterRotOffset, AnglesVec,
                 ObjSize, OS_number=None, quant=True):
        self.OS_number = 1 if OS_number is None else OS_number
        self.objsize_user_given = None if DetectorsDimH_pad == 0 else ObjSize
        if DetectorsDimH_pad > 0:
            ObjSize = DetectorsDimH + 2 * DetectorsDimH_pad
        if DetectorsDimV == 0 or DetectorsDimV is None:
            DetectorsDimV = 1
        self.Atools = Atools(DetectorsDimH, DetectorsDimH_pad, DetectorsDimV, AnglesVec,
                             CenterRotOffset, ObjSize, OS_number, quant)

    # methodsIR_CuPy.py:116-126
    def _Ax(self, x, sub_ind=1, os=False):
        return self.Atools._forwprojOSCuPy(x, sub_ind) if os else self.Atools._forwprojCuPy(x)

    def _Atb(self, b, sub_ind=1, os=False):
        return self.Atools._backprojOSCuPy(b, sub_ind) if os else self.Atools._backprojCuPy(b)

    def _finish(self, x, mask_radius):
        if self.objsize_user_given is not None:
            return recon_crop(x, self.objsize_user_given)
        if mask_radius is not None:
            circular_mask(x, mask_radius)
        return x

    def _prep(self, data):
        data = np.asarray(data, dtype=np.float32)
        if data.ndim == 2:
            data = data[None]
        return pad_detector(data, self.Atools.detectors_x_pad)

    def _subset(self, s):
        return subset_indices(self.Atools.newInd_Vec, self.Atools.NumbProjBins, s)

    # data_fidelities.py:7-40
    def grad_data_term(self, x, b, use_os, sub_ind, indVec, w, fidelity="LS"):
        if fidelity in ("LS", "PWLS"):
            res = self._Ax(x, sub_ind, use_os) - b
            if w is not None:
                res *= w[:, indVec, :] if use_os else w
        else:  # KL
            res = 1 - b / np.clip(self._Ax(x, sub_ind, use_os), 1e-8, None)
            res = res.astype(np.float32)
        return self._Atb(res, sub_ind, use_os)

    # methodsIR_CuPy.py:128-172
    def Landweber(self, data, iterations=1500, tau=1e-5, nonneg=False, mask_radius=1.0):
        b = self._prep(data)
        x = np.zeros(self.Atools.vol_shape, dtype=np.float32)
        for _ in range(iterations):
            x -= f32(tau) * self._Atb(self._Ax(x) - b)
            if nonneg:
                np.maximum(x, 0, out=x)
        return self._finish(x, mask_radius)

    # methodsIR_CuPy.py:174-231
    def SIRT(self, data, iterations=200, nonneg=False, mask_radius=1.0):
        b = self._prep(data)
        with np.errstate(divide="ignore", invalid="ignore"):
            R = 1.0 / self._Ax(np.ones(self.Atools.vol_shape, dtype=np.float32))
            R = np.nan_to_num(R, copy=False, nan=1.0, posinf=1.0, neginf=1.0)
            C = 1.0 / self._Atb(np.ones(b.shape, dtype=np.float32))
            C = np.nan_to_num(C, copy=False, nan=1.0, posinf=1.0, neginf=1.0)
        x = np.ones(self.Atools.vol_shape, dtype=np.float32)
        for _ in range(iterations):
            x += C * self._Atb(R * (b - self._Ax(x)))
            if nonneg:
                np.maximum(x, 0, out=x)
        return self._finish(x, mask_radius)

    # methodsIR_CuPy.py:233-309 (with a contiguous first back-projection, i.e. without the
    # swapaxes-view bug of SURVEY.md section 0 item 2)
    def CGLS(self, data, iterations=30, nonneg=False, mask_radius=1.0, first_bp_data=None):
        """``first_bp_data``: what the FIRST back-projection reads instead of ``data`` -- the reference hands ASTRA
        the raw buffer of an axis-swapped view there (methodsIR_CuPy.py:270 with astra_base.py:533-535), and its
        goldens tests/test_RecToolsIRCuPy.py:152-153, 216-217 encode that."""
        b = self._prep(data)
        shp = self.Atools.vol_shape
        x = np.zeros(int(np.prod(shp)), dtype=np.float32)
        d = self._Atb(b if first_bp_data is None else self._prep(first_bp_data)).ravel()
        normr2 = np.inner(d, d)
        r = b.copy().ravel()
        for _ in range(iterations):
            Ad = self._Ax(d.reshape(shp)).ravel()
            alpha = normr2 / np.inner(Ad, Ad)
            x += alpha * d
            r -= alpha * Ad
            s = self._Atb(r.reshape(b.shape)).ravel()
            normr2_new = np.inner(s, s)
            beta = normr2_new / normr2
            normr2 = normr2_new
            d = s + beta * d
            if nonneg:
                np.maximum(x, 0, out=x)
        return self._finish(x.reshape(shp), mask_radius)

    # methodsIR_CuPy.py:311-354
    def powermethod(self, seed=0, iterations=15):
        rng = np.random.default_rng(seed)
        x1 = rng.standard_normal(self.Atools.vol_shape).astype(np.float32)
        os_mode = self.OS_number > 1
        y = self._Ax(x1, 0, os_mode)
        s = 1.0
        for _ in range(iterations):
            x1 = self._Atb(y, 0, os_mode)
            s = np.linalg.norm(x1.ravel())
            x1 = x1 / s
            y = self._Ax(x1, 0, os_mode)
        return float(s)

    def _init(self, data, fidelity, initialise, lipschitz_const, ones=False):
        b = self._prep(data)
        L = self.powermethod() if lipschitz_const is None else lipschitz_const
        if initialise is not None and initialise.shape == self.Atools.vol_shape:
            x0 = np.array(initialise, dtype=np.float32)
        else:
            x0 = (np.ones if ones else np.zeros)(self.Atools.vol_shape, dtype=np.float32)
        w = None
        if fidelity == "PWLS":  # methodsIR_CuPy.py:392-395
            w = np.maximum(b, f32(1e-6))
            w = w / w.max()
        return b, L, x0, w

    # Extension (no code in this reference snapshot; legacy call sites
    # Demos/methods_IR_legacy/DemoFISTA_artifacts2D.py:197,243,307-309): residual of the Huber /
    # Group-Huber ring / SWLS data terms.  Defines the semantics the CUDA path (k_resid_post) is held to.
    def residual_ext(self, x, b, use_os, sub_ind, indVec, w, fidelity, huber, r_x, alpha, beta, studentst=None):
        res = (self._Ax(x, sub_ind, use_os) - b).astype(np.float32)
        vec = None
        if r_x is not None:
            res = res + f32(alpha) * r_x[:, None, :]
            vec = np.zeros_like(r_x)
            for a in range(res.shape[1]):  # sequential fp32 sum over the angles
                vec += res[:, a, :]
        if huber is not None:
            absr = np.abs(res)
            with np.errstate(divide="ignore", invalid="ignore"):
                res = np.where(absr > f32(huber), res * (f32(huber) / absr), res).astype(np.float32)
        if studentst is not None:  # Student's-t penalty log(1 + r^2 / sigma^2): gradient 2 r / (sigma^2 + r^2)
            res = ((f32(2.0) * res) / (f32(f32(studentst) * f32(studentst)) + res * res)).astype(np.float32)
        if fidelity in ("PWLS", "SWLS"):
            ws = w[:, indVec, :] if use_os else w
            res = res * ws
            if fidelity == "SWLS":
                s = np.zeros((res.shape[0], res.shape[2]), np.float32)
                sw = np.zeros_like(s)
                for a in range(res.shape[1]):
                    s += res[:, a, :]
                    sw += ws[:, a, :]
                res = res - ws * (s / (sw + f32(beta)))[:, None, :]
        return res.astype(np.float32), vec

    # methodsIR_CuPy.py:401-484
    def FISTA(self, data, iterations, lipschitz_const=None, regularisation=None, nonneg=False,
              fidelity="LS", initialise=None, mask_radius=1.0, huber_threshold=None, ringGH_lambda=None,
              ringGH_accelerate=50, beta_SWLS=0.1, studentst_threshold=None):
        reg = _reg_defaults(regularisation)
        b_all, L, x0, w = self._init(data, "PWLS" if fidelity == "SWLS" else fidelity, initialise, lipschitz_const)
        extended = (huber_threshold is not None or ringGH_lambda is not None or fidelity == "SWLS"
                    or studentst_threshold is not None)
        r = r_x = None
        if ringGH_lambda is not None:
            r = np.zeros((b_all.shape[0], b_all.shape[2]), np.float32)
            r_x = r.copy()
        use_os = self.OS_number > 1
        Linv = 1.0 / L
        b = b_all
        indVec = None
        t = f32(1.0)
        X_t = x0.copy()
        X = x0.copy()
        for _ in range(iterations):
            for sub in range(self.OS_number):
                X_old = X
                t_old = t
                if use_os:
                    indVec = self._subset(sub)
                    b = b_all[:, indVec, :]
                if extended:
                    res, vec = self.residual_ext(X_t, b, use_os, sub, indVec, w, fidelity, huber_threshold, r_x,
                                                 ringGH_accelerate, beta_SWLS, studentst_threshold)
                    grad = self._Atb(res, sub, use_os)
                    if r is not None:
                        r_old = r
                        r = (r_x - f32(Linv) * vec).astype(np.float32)
                else:
                    grad = self.grad_data_term(X_t, b, use_os, sub, indVec, w, fidelity)
                X = (X_t - Linv * grad).astype(np.float32)
                if nonneg:
                    np.maximum(X, 0, out=X)
                if reg["method"] is not None:
                    X = prox_regul(X, reg, 1 if nonneg else 0)
                t = f32((1.0 + np.sqrt(1.0 + 4.0 * t ** 2)) * 0.5)
                coef = f32((t_old - 1.0) / t)
                X_t = (X + coef * (X - X_old)).astype(np.float32)
                if r is not None:
                    r = (np.maximum(np.abs(r) - f32(ringGH_lambda), 0) * np.sign(r)).astype(np.float32)
                    r_x = (r + coef * (r - r_old)).astype(np.float32)
        return self._finish(X, mask_radius)

    # methodsIR_CuPy.py:486-585
    def ADMM(self, data, iterations, lipschitz_const=None, regularisation=None, nonneg=False,
             fidelity="LS", initialise=None, rho=1.0, relax=1.6, mask_radius=1.0):
        reg = _reg_defaults(regularisation)
        b_all, L, x0, w = self._init(data, fidelity, initialise, lipschitz_const)
        use_os = self.OS_number > 1
        b = b_all
        indVec = None
        x = x0.copy()
        z = x0.copy()
        z_old = 0
        u = np.zeros_like(x0)
        tau = 0.9 / (L + rho)
        reg["regul_param"] = reg["regul_param"] / rho
        for it in range(iterations):
            for sub in range(self.OS_number):
                if use_os:
                    indVec = self._subset(sub)
                    b = b_all[:, indVec, :]
                grad_data = self.grad_data_term(z, b, use_os, sub, indVec, w, fidelity)
                grad_admm = rho * (z - x + u)
                z = (z - tau * (grad_data + grad_admm)).astype(np.float32)
                if nonneg:
                    np.maximum(z, 0, out=z)
                if it > 1:
                    z = ((1.0 - relax) * z_old + relax * z).astype(np.float32)
                z_old = z.copy()
                xp = z + u
                x = prox_regul(xp, reg, 1 if nonneg else 0) if reg["method"] is not None else xp
            u = u + (z - x)
        return self._finish(x, mask_radius)

    # methodsIR_CuPy.py:587-667
    def OSEM(self, data, iterations, regularisation=None, mask_radius=1.0):
        reg = _reg_defaults(regularisation)
        b_all, _, x, _ = self._init(data, "KL", None, 0.0, ones=True)
        use_os = self.OS_number > 1
        eps = 1e-8
        b = b_all
        if not use_os:
            norm = self._Atb(np.ones_like(b))
        else:
            norm = self._Atb(np.ones_like(b_all[:, self._subset(0), :]), 0, True)
        norm = np.clip(norm, eps, None)
        for _ in range(iterations):
            for sub in range(self.OS_number):
                if use_os:
                    b = b_all[:, self._subset(sub), :]
                Ax = np.clip(self._Ax(x, sub, use_os), eps, None)
                x = x * (self._Atb((b / Ax).astype(np.float32), sub, use_os) * norm)
                if reg["method"] is not None:
                    x = prox_regul(x.astype(np.float32), reg, 0)
        return self._finish(x, mask_radius)


class RecDIR:
    """Oracle of ``RecToolsDIRCuPy`` FORWPROJ/BACKPROJ/FBP (methodsDIR_CuPy.py:70-150)."""

    def __init__(self, DetectorsDimH, DetectorsDimH_pad, DetectorsDimV, CenterRotOffset, AnglesVec,
                 ObjSize, quant=True):
        if DetectorsDimV == 0 or DetectorsDimV is None:
            DetectorsDimV = 1
        self.Atools = Atools(DetectorsDimH, DetectorsDimH_pad, DetectorsDimV, AnglesVec,
                             CenterRotOffset, ObjSize, None, quant)

    def FORWPROJ(self, vol):
        return self.Atools._forwprojCuPy(vol)

    def BACKPROJ(self, data):
        """data [detY, angles, detX] (contiguous)."""
        return self.Atools._backprojCuPy(pad_detector(np.asarray(data, np.float32),
                                                      self.Atools.detectors_x_pad))

    def FBP(self, data, cutoff_freq=0.35, recon_mask_radius=None):
        """data [angles, detY, detX] (methodsDIR_CuPy.py:114-150)."""
        data = pad_detector(np.asarray(data, dtype=np.float32), self.Atools.detectors_x_pad)
        data = filtersinc3d(data, cutoff_freq)
        data = np.ascontiguousarray(np.swapaxes(data, 0, 1))
        rec = self.Atools._backprojCuPy(data)
        if recon_mask_radius is not None:
            circular_mask(rec, recon_mask_radius)
        return rec
