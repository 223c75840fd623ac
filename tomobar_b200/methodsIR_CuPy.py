"""Iterative reconstruction on B200 behind the ``RecToolsIRCuPy`` interface.

Public surface (constructor, method names, the three parameter dictionaries, return shapes)
follows tomobar/methodsIR_CuPy.py:36-667.  The loops are restructured around libtmb.so:

* ``grad_data_term`` is one fused call (forward projector with the residual / weighting in its
  epilogue, written straight into the back-projector's layout, then the back-projector);
* the FISTA / ADMM elementwise updates are single fused kernels over preallocated volumes;
* the TV prox runs in a preallocated output buffer.

Arrays are float32 CUDA torch tensors (CuPy / numpy inputs are converted).  There is no CPU
path.  Optional z-sharding over ``torch.distributed`` ranks is provided by
``tomobar_b200.zshard``.
"""

from __future__ import annotations

import math
from typing import Optional, Union

import numpy as np
import torch

from tomobar_b200._lib import lib, check
from tomobar_b200._tensors import as_cuda_f32, ptr, stream_ptr
from tomobar_b200.projector import ProjTools3D
from tomobar_b200.regularisersCuPy import PD_TV_cupy, ROF_TV_cupy, prox_regul
from tomobar_b200.supp.dicts import dicts_check
from tomobar_b200.supp.funcs import _raw_buffer_view
from tomobar_b200.supp.suppTools import _apply_horiz_detector_padding, check_kwargs, perform_recon_crop


def _count(t: torch.Tensor) -> int:
    return t.numel()


class RecToolsIRCuPy:
    """Iterative reconstruction algorithms (FISTA, ADMM, Landweber, SIRT, CGLS, OSEM) with the
    projector pair, the data-fidelity gradient and the TV proximal operators running as
    hand-written sm_100a kernels.

    Args (methodsIR_CuPy.py:53-69):
        DetectorsDimH (int): Horizontal detector dimension size.
        DetectorsDimH_pad (int): The amount of padding for the horizontal detector.
        DetectorsDimV (int, None): Vertical detector dimension size, 'None' for 2D.
        CenterRotOffset (float, np.ndarray): Centre of Rotation scalar or one value per angle.
        AnglesVec (np.ndarray): Projection angles in radians.
        ObjSize (int): Size of the reconstructed slice [ObjSize, ObjSize].
        device_projector (int): GPU index.
        OS_number (int, None): number of ordered subsets, None for non-OS reconstruction.
    """

    def __init__(
        self,
        DetectorsDimH: int,
        DetectorsDimH_pad: int,
        DetectorsDimV: Union[int, None],
        CenterRotOffset: Union[float, np.ndarray],
        AnglesVec: np.ndarray,
        ObjSize: int,
        device_projector: int = 0,
        OS_number: Optional[int] = None,
        quantise_weights: bool = True,
    ):
        self.OS_number = OS_number
        self.objsize_user_given = None if DetectorsDimH_pad == 0 else ObjSize
        if DetectorsDimH_pad > 0:
            # padded detector => reconstruct on the padded grid, crop at the end (:77-79)
            ObjSize = DetectorsDimH + 2 * DetectorsDimH_pad
        if DetectorsDimV == 0 or DetectorsDimV is None:
            DetectorsDimV = 1  # 2-D is a one-slice 3-D problem (:81-82)
        self.geom = "3D"
        self.Atools = ProjTools3D(
            DetectorsDimH,
            DetectorsDimH_pad,
            DetectorsDimV,
            AnglesVec,
            CenterRotOffset,
            ObjSize,
            "gpu",
            device_projector,
            OS_number,
            quantise_weights=quantise_weights,
        )
        self.data_fidelity = "LS"
        self.nonneg_regul = 0
        self.power_seed = 0  # the reference draws an unseeded cp.random.randn (:326)
        # True: CGLS's first back-projection reproduces the reference's non-contiguous-view behaviour (:270 with
        # astra_base.py:533-535) -- the goldens tests/test_RecToolsIRCuPy.py:152-153,
        # 216-217 encode it
        self.compat_view_bug = False
        self.zshard = None   # set_zshard(): this object reconstructs one z-block of a larger volume
        self.tv_peer_memory = None  # sharded TV halos: None = NVLink peer loads on NCCL, False = messages
        self.tv_sync = "signals"    # peer-memory ordering: pairwise semaphores, or "barrier"
        self.tv_pairs = None        # sharded PD_TV: two iterations per pass (None: on unless TMB_SHARDED_PAIRS=0)
        self._sharded_tv = {}

    def set_zshard(self, shard) -> None:
        """Declare that this object holds the slices ``[shard.z0, shard.z1)`` of a volume that is
        z-sharded over ``torch.distributed`` ranks (``tomobar_b200.zshard.ZShard``).  The norms of
        the power method and of CGLS, the PWLS weight normalisation and the 3-D PD_TV prox then act
        on the whole volume (scalar all-reduces, one-plane halo exchange per inner TV iteration)."""
        if shard is not None and shard.nz_local != self.Atools.detectors_y:
            raise ValueError("set_zshard: DetectorsDimV must equal the number of slices of the shard")
        self.zshard = shard
        self._sharded_tv = {}

    @property
    def OS_number(self) -> int:
        return self._OS_number

    @OS_number.setter
    def OS_number(self, value):
        self._OS_number = 1 if value is None else value

    @property
    def objsize_user_given(self):
        return self._objsize_user_given

    @objsize_user_given.setter
    def objsize_user_given(self, value):
        self._objsize_user_given = value

    # ---- operator shorthands (:116-126) ---------------------------------------------------------
    def _Ax(self, x, sub_ind: int = 1, os: bool = False):
        return self.Atools._forwprojOSCuPy(x, os_index=sub_ind) if os else self.Atools._forwprojCuPy(x)

    def _Atb(self, b, sub_ind: int = 1, os: bool = False):
        return self.Atools._backprojOSCuPy(b, os_index=sub_ind) if os else self.Atools._backprojCuPy(b)

    # ---- helpers --------------------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.Atools.device).cuda_stream

    def _zeros_vol(self):
        return torch.zeros(self.Atools.vol_geom, dtype=torch.float32, device=self.Atools.device)

    def _prepare_data(self, _data_upd_: dict) -> torch.Tensor:
        data = _apply_horiz_detector_padding(_data_upd_["projection_data"], self.Atools.detectors_x_pad, True)
        data = data.contiguous()
        _data_upd_["projection_data"] = data
        return data

    def _finish(self, x: torch.Tensor, recon_mask_radius):
        if self.objsize_user_given is not None:
            return perform_recon_crop(x, self.objsize_user_given)
        return check_kwargs(x, cupyrun=True, recon_mask_radius=recon_mask_radius)

    def _axpy(self, a: float, x: torch.Tensor, y: torch.Tensor, nonneg: bool = False) -> None:
        check(lib.tmb_axpy(float(a), ptr(x), ptr(y), _count(y), int(nonneg), self._stream()), "tmb_axpy")

    def _norm(self, x: torch.Tensor) -> torch.Tensor:
        if self.zshard is not None and self.zshard.world > 1:
            return self.zshard.norm(x)
        return torch.linalg.vector_norm(x.ravel())

    def _dot(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        if self.zshard is not None and self.zshard.world > 1:
            return self.zshard.dot(a, b)
        return torch.inner(a.ravel(), b.ravel())

    def _subset_indices(self, sub_ind: int) -> np.ndarray:
        indVec = self.Atools.newInd_Vec[sub_ind, :]
        if indVec[self.Atools.NumbProjBins - 1] == 0:
            indVec = indVec[:-1]
        return indVec

    # ---- simple algorithms ------------------------------------------------------------------------
    def Landweber(self, _data_: dict, _algorithm_: Union[dict, None] = None) -> torch.Tensor:
        """x <- x - tau * A^T(Ax - b)   (:128-172)"""
        _data_upd_, _algorithm_upd_, _ = dicts_check(self, _data_, _algorithm_, method_run="Landweber")
        b = self._prepare_data(_data_upd_)
        x_rec = self._zeros_vol()
        grad = torch.empty_like(x_rec)
        tau = _algorithm_upd_["tau_step_lanweber"]
        with torch.cuda.device(self.Atools.device):
            for _ in range(_algorithm_upd_["iterations"]):
                self.Atools.grad_data_term(x_rec, b, None, "LS", None, out=grad)
                self._axpy(-tau, grad, x_rec, _algorithm_upd_["nonnegativity"])
        return self._finish(x_rec, _algorithm_upd_["recon_mask_radius"])

    def SIRT(self, _data_: dict, _algorithm_: Union[dict, None] = None) -> torch.Tensor:
        """x <- x + C A^T R (b - Ax)   (:174-231)"""
        _data_upd_, _algorithm_upd_, _ = dicts_check(self, _data_, _algorithm_, method_run="SIRT")
        b = self._prepare_data(_data_upd_)
        A = self.Atools
        R = 1.0 / A._forwprojCuPy(torch.ones(A.vol_geom, dtype=torch.float32, device=A.device))
        R = torch.nan_to_num(R, nan=1.0, posinf=1.0, neginf=1.0)
        C = 1.0 / A._backprojCuPy(torch.ones(A.proj_geom, dtype=torch.float32, device=A.device))
        C = torch.nan_to_num(C, nan=1.0, posinf=1.0, neginf=1.0)
        x_rec = torch.ones(A.vol_geom, dtype=torch.float32, device=A.device)
        for _ in range(_algorithm_upd_["iterations"]):
            x_rec += C * A._backprojCuPy(R * (b - A._forwprojCuPy(x_rec)))
            if _algorithm_upd_["nonnegativity"]:
                x_rec.clamp_(min=0)
        return self._finish(x_rec, _algorithm_upd_["recon_mask_radius"])

    def CGLS(self, _data_: dict, _algorithm_: Union[dict, None] = None) -> torch.Tensor:
        """Conjugate gradients on the normal equations (:233-309).  The first back-projection
        uses the logical (contiguous) data; the reference passes a possibly strided view's raw
        pointer there."""
        _data_upd_, _algorithm_upd_, _ = dicts_check(self, _data_, _algorithm_, method_run="CGLS")
        first = None
        if self.compat_view_bug and self.Atools.detectors_x_pad == 0:
            first = _raw_buffer_view(_data_upd_["projection_data"]).contiguous()
        b = self._prepare_data(_data_upd_)
        A = self.Atools
        x_rec = self._zeros_vol()
        d = A._backprojCuPy(b if first is None else first)
        normr2 = self._dot(d, d)
        r = b.clone()
        for _ in range(_algorithm_upd_["iterations"]):
            Ad = A._forwprojCuPy(d)
            alpha = normr2 / self._dot(Ad, Ad)
            x_rec += alpha * d
            r -= alpha * Ad
            s = A._backprojCuPy(r)
            normr2_new = self._dot(s, s)
            beta = normr2_new / normr2
            normr2 = normr2_new
            d = s + beta * d
            if _algorithm_upd_["nonnegativity"]:
                x_rec.clamp_(min=0)
        return self._finish(x_rec, _algorithm_upd_["recon_mask_radius"])

    def powermethod(self, _data_: dict) -> float:
        """Largest eigenvalue of A^T A by 15 power iterations (:311-354); subset 0 only in OS mode.
        The PWLS branch of the reference multiplies by a ones array and is a no-op."""
        if _data_.get("data_fidelity") is None:
            _data_["data_fidelity"] = "LS"
        A = self.Atools
        gen = torch.Generator(device=A.device)
        gen.manual_seed(self.power_seed + (self.zshard.z0 if self.zshard is not None else 0))
        x1 = torch.randn(A.vol_geom, dtype=torch.float32, device=A.device, generator=gen)
        sub = 0 if self.OS_number > 1 else None
        s = 1.0
        y = A._forwprojOSCuPy(x1, 0) if sub is not None else A._forwprojCuPy(x1)
        for _ in range(15):
            x1 = A._backprojOSCuPy(y, 0) if sub is not None else A._backprojCuPy(y)
            s = self._norm(x1)
            x1 = x1 / s
            y = A._forwprojOSCuPy(x1, 0) if sub is not None else A._forwprojCuPy(x1)
        return float(s)

    def _common_initialisation(self, _data_, _algorithm_, _regularisation_, method_run):
        """(:356-399)"""
        _data_upd_, _algorithm_upd_, _regularisation_upd_ = dicts_check(
            self, _data_, _algorithm_, _regularisation_, method_run=method_run
        )
        b = self._prepare_data(_data_upd_)
        if _algorithm_upd_.get("lipschitz_const") is None:
            _algorithm_upd_["lipschitz_const"] = self.powermethod(_data_upd_)

        rec_dim = self.Atools.vol_geom
        init = _algorithm_upd_["initialise"]
        fill = 1.0 if method_run == "OSEM" else 0.0
        if init is not None:
            init = as_cuda_f32(init, self.Atools.device, "initialisation")
            if tuple(init.shape) == tuple(rec_dim):
                x0 = init.contiguous()
            else:
                print(
                    f"Provided initialisation (array) has incorrect dimensions, the correct dims are {rec_dim}. "
                    "Zero initialisation is used."
                )
                x0 = torch.zeros(rec_dim, dtype=torch.float32, device=self.Atools.device)
        else:
            x0 = torch.full(rec_dim, fill, dtype=torch.float32, device=self.Atools.device)

        use_os = self.OS_number > 1
        w = None
        if _data_["data_fidelity"] in ["PWLS", "SWLS"]:
            w = torch.clamp(b, min=1e-6)  # weights for the PWLS model (:392-395)
            wmax = w.max() if self.zshard is None else self.zshard.max(w)
            w = w / wmax
        return (_data_upd_, _algorithm_upd_, _regularisation_upd_, x0, w, use_os)

    def _prox_into(self, X: torch.Tensor, reg: dict, out: torch.Tensor) -> torch.Tensor:
        """``prox_regul`` (regularisersCuPy.py:6-38) writing into a preallocated volume."""
        dev = self.Atools.device_index
        sh = self.zshard
        # the same decision on every rank: taken from the GLOBAL slice count and the in-plane shape, never from
        # this rank's block (a one-slice last block must not send one rank down the local 2-D path while its
        # neighbours wait for it; ZShard.require_tv_shards raises on all ranks instead)
        sharded3d = sh is not None and sh.world > 1 and X.ndim == 3 and sh.nz_total > 1 and min(X.shape[1:]) > 1
        if "ROF_TV" in reg["method"]:
            if sharded3d:
                from tomobar_b200.zshard import ShardedROFTV

                key = ("rof", tuple(X.shape), bool(reg.get("half_precision", False)))
                if key not in self._sharded_tv:
                    self._sharded_tv = {key: ShardedROFTV(sh, key[1], X.device, key[2], self.tv_peer_memory, self.tv_sync)}
                return self._sharded_tv[key](X, reg["regul_param"], reg["iterations"], reg["time_marching_step"],
                                             out=out)
            return ROF_TV_cupy(X, reg["regul_param"], reg["iterations"], reg["time_marching_step"], dev,
                               reg.get("half_precision", False), out=out)
        if "PD_TV" in reg["method"]:
            if sharded3d:
                # whole-volume 3-D TV across the z-shards: halo exchange between inner iterations
                from tomobar_b200.zshard import ShardedPDTV

                key = ("pd", tuple(X.shape), bool(reg.get("half_precision", False)))
                if key not in self._sharded_tv:
                    self._sharded_tv = {key: ShardedPDTV(sh, key[1], X.device, key[2], self.tv_peer_memory, self.tv_sync,
                                                          self.tv_pairs)}
                return self._sharded_tv[key](X, reg["regul_param"], reg["iterations"], reg["methodTV"],
                                             self.nonneg_regul, reg["PD_LipschitzConstant"], out=out)
            return PD_TV_cupy(X, reg["regul_param"], reg["iterations"], reg["methodTV"], self.nonneg_regul,
                              reg["PD_LipschitzConstant"], dev, reg.get("half_precision", False), out=out)
        raise ValueError(f"Unknown regularisation method {reg['method']!r}: ROF_TV and PD_TV are supported")

    # ---- FISTA (:401-484) -------------------------------------------------------------------------
    def FISTA(
        self,
        _data_: dict,
        _algorithm_: Union[dict, None] = None,
        _regularisation_: Union[dict, None] = None,
    ) -> torch.Tensor:
        """Ordered-subsets FISTA with LS / PWLS / KL data terms and ROF_TV / PD_TV proximal steps.

        Per sub-step: grad = A_s^T(W(A_s X_t - b_s));  X = X_t - grad/L;  [X = max(X, 0)];
        [X = prox(X)];  t = (1 + sqrt(1 + 4 t^2))/2 (float32);  X_t = X + ((t_old - 1)/t)(X - X_old).
        """
        (_data_upd_, _algorithm_upd_, _regularisation_upd_, x0, w, use_os) = self._common_initialisation(
            _data_, _algorithm_, _regularisation_, method_run="FISTA"
        )
        A = self.Atools
        b = _data_upd_["projection_data"]
        L_const_inv = 1.0 / _algorithm_upd_["lipschitz_const"]
        nonneg = bool(_algorithm_upd_["nonnegativity"])
        regularised = _regularisation_upd_["method"] is not None
        st = self._stream()
        count = x0.numel()

        t = np.float32(1.0)
        X_t = x0.clone()
        X = x0.clone()
        X_old = torch.empty_like(x0)   # rotating volumes: X_old / X / scratch
        G = torch.empty_like(x0)       # gradient, then the pre-prox iterate

        # robust / ring-artefact data terms (extension): Huber residual clipping, the
        # Group-Huber ring model (one offset r per detector pixel, soft-thresholded, with its own
        # momentum) and stripe-weighted least squares
        huber = _data_upd_.get("huber_threshold")
        studentst = _data_upd_.get("studentst_threshold")
        ring_lambda = _data_upd_.get("ringGH_lambda")
        extended = huber is not None or studentst is not None or ring_lambda is not None or self.data_fidelity == "SWLS"
        if huber is not None and studentst is not None:
            raise ValueError("huber_threshold and studentst_threshold exclude each other")
        if extended and self.data_fidelity == "KL":
            raise ValueError("Huber / ring / SWLS models combine with the LS and PWLS data terms only")
        r = r_x = vec = None
        if ring_lambda is not None:
            r = torch.zeros((A.detectors_y, A.nu), dtype=torch.float32, device=A.device)
            r_x, vec = r.clone(), torch.empty_like(r)

        with torch.cuda.device(A.device):
            for _ in range(_algorithm_upd_["iterations"]):
                for sub_ind in range(self.OS_number):
                    X_old, X = X, X_old            # X_old <- current iterate; X <- free buffer
                    t_old = t
                    if extended:
                        A.grad_data_term_ext(X_t, b, sub_ind if use_os else None, self.data_fidelity, w, huber,
                                             r_x, float(_data_upd_["ringGH_accelerate"]),
                                             float(_data_upd_["beta_SWLS"]), vec, out=G, studentst_threshold=studentst)
                        if r is not None:
                            r_old = r
                            r = r_x - np.float32(L_const_inv) * vec
                    else:
                        A.grad_data_term(X_t, b, sub_ind if use_os else None, self.data_fidelity, w, out=G)
                    target = G if regularised else X
                    check(lib.tmb_fista_grad_step(ptr(X_t), ptr(G), ptr(target), count, float(L_const_inv),
                                                  int(nonneg), st), "tmb_fista_grad_step")
                    if regularised:
                        self._prox_into(G, _regularisation_upd_, X)
                    t = np.float32((1.0 + np.sqrt(1.0 + 4.0 * t**2)) * 0.5)
                    coef = np.float32((t_old - 1.0) / t)
                    check(lib.tmb_fista_momentum(ptr(X), ptr(X_old), ptr(X_t), count, float(coef), st),
                          "tmb_fista_momentum")
                    if r is not None:
                        r = torch.clamp(r.abs() - np.float32(ring_lambda), min=0) * torch.sign(r)
                        r_x = (r + coef * (r - r_old)).contiguous()
        return self._finish(X, _algorithm_upd_["recon_mask_radius"])

    # ---- ADMM (:486-585) --------------------------------------------------------------------------
    def ADMM(
        self,
        _data_: dict,
        _algorithm_: Union[dict, None] = None,
        _regularisation_: Union[dict, None] = None,
    ) -> torch.Tensor:
        """Linearised, relaxed ADMM: tau = 0.9/(L + rho); the regularisation parameter is divided
        by rho in the caller's dictionary (like the reference, :526-528); relaxation starts at the
        third outer iteration (``iter_no > 1``, :551)."""
        (_data_upd_, _algorithm_upd_, _regularisation_upd_, x0, w, use_os) = self._common_initialisation(
            _data_, _algorithm_, _regularisation_, method_run="ADMM"
        )
        A = self.Atools
        b = _data_upd_["projection_data"]
        rho = _algorithm_upd_["ADMM_rho_const"]
        alpha = _algorithm_upd_["ADMM_relax_par"]
        nonneg = bool(_algorithm_upd_["nonnegativity"])
        regularised = _regularisation_upd_["method"] is not None
        st = self._stream()
        count = x0.numel()

        x = x0.clone()
        z = x0.clone()
        z_old = torch.zeros_like(x0)
        u = torch.zeros_like(x0)
        grad = torch.empty_like(x0)
        xprox = torch.empty_like(x0)

        tau = 0.9 / (_algorithm_upd_["lipschitz_const"] + rho)
        _regularisation_upd_["regul_param"] = _regularisation_upd_["regul_param"] / rho

        with torch.cuda.device(A.device):
            for iter_no in range(_algorithm_upd_["iterations"]):
                for sub_ind in range(self.OS_number):
                    A.grad_data_term(z, b, sub_ind if use_os else None, self.data_fidelity, w, out=grad)
                    target = xprox if regularised else x
                    check(lib.tmb_admm_z_step(ptr(z), ptr(z_old), ptr(x), ptr(u), ptr(grad), ptr(target), count,
                                              float(tau), float(rho), int(nonneg), int(iter_no > 1), float(alpha),
                                              st), "tmb_admm_z_step")
                    if regularised:
                        self._prox_into(xprox, _regularisation_upd_, x)
                check(lib.tmb_admm_u_step(ptr(u), ptr(z), ptr(x), count, st), "tmb_admm_u_step")
                if _algorithm_upd_["verbose"]:
                    if np.mod(iter_no, (round)(_algorithm_upd_["iterations"] / 5) + 1) == 0:
                        print("ADMM iteration (", iter_no + 1, ") using", _regularisation_upd_["method"],
                              "regularisation")
        return self._finish(x, _algorithm_upd_["recon_mask_radius"])

    # ---- OSEM (:587-667) --------------------------------------------------------------------------
    def OSEM(
        self,
        _data_: dict,
        _algorithm_: Union[dict, None] = None,
        _regularisation_: Union[dict, None] = None,
    ) -> torch.Tensor:
        """Ordered-subsets expectation maximisation (MLEM when OS_number is None)."""
        (_data_upd_, _algorithm_upd_, _regularisation_upd_, x, w, use_os) = self._common_initialisation(
            _data_, _algorithm_, _regularisation_, method_run="OSEM"
        )
        eps = 1e-8
        b_all = _data_upd_["projection_data"]
        proj_data = b_all
        if not use_os:
            normalisation = self._Atb(torch.ones_like(b_all))
        else:
            ind0 = torch.as_tensor(self._subset_indices(0), device=b_all.device)
            normalisation = self._Atb(torch.ones_like(b_all[:, ind0, :]), 0, use_os)
        normalisation = torch.clamp(normalisation, min=eps)

        for _ in range(_algorithm_upd_["iterations"]):
            for sub_ind in range(self.OS_number):
                if use_os:
                    ind = torch.as_tensor(self._subset_indices(sub_ind), device=b_all.device)
                    proj_data = b_all[:, ind, :]
                Ax = torch.clamp(self._Ax(x, sub_ind, use_os), min=eps)
                backproj = self._Atb(proj_data / Ax, sub_ind, use_os)
                x = x * (backproj * normalisation)
                if _regularisation_upd_["method"] is not None:
                    # through _prox_into: whole-volume TV also when this object holds one z-shard
                    x = self._prox_into(x.contiguous(), _regularisation_upd_, torch.empty_like(x))
        return self._finish(x, _algorithm_upd_["recon_mask_radius"])
