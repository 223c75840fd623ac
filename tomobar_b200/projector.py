"""Operator boundary: a drop-in for ``AstraTools3D`` (astra_wrappers/astra_tools3d.py:19-110)
whose forward/back projections run in libtmb.so instead of astra-toolbox.

The four methods ``_forwprojCuPy`` / ``_forwprojOSCuPy`` / ``_backprojCuPy`` /
``_backprojOSCuPy`` keep the reference names, shapes and ownership rules (inputs untouched,
fresh output per call) so ``methodsIR_CuPy.py`` / ``methodsDIR_CuPy.py`` can use this object
as their ``Atools``.  Arrays are float32 CUDA torch tensors (CuPy arrays are accepted through
DLPack / ``__cuda_array_interface__``).
"""

from __future__ import annotations

import ctypes as C
from typing import Optional, Union

import numpy as np
import torch

from tomobar_b200._lib import lib, check
from tomobar_b200._tensors import as_cuda_f32, ptr, stream_ptr, require_dense

FIDELITY = {"LS": 0, "PWLS": 1, "KL": 2}


class ProjTools3D:
    """3-D parallel-beam projector pair (vertical rotation axis, unit pixels).

    Args mirror ``AstraTools3D.__init__`` (astra_tools3d.py:26-38); validation mirrors the
    property setters of ``AstraBase`` (astra_base.py:74-193).
    """

    def __init__(
        self,
        detectors_x: int,
        detectors_x_pad: int,
        detectors_y: int,
        angles_vec: np.ndarray,
        centre_of_rotation: Union[float, np.ndarray, None],
        recon_size: int,
        processing_arch: str = "gpu",
        device_index: int = 0,
        ordsub_number: Optional[int] = None,
        verbosity: bool = False,
        quantise_weights: bool = True,
    ):
        if detectors_x <= 0:
            raise ValueError("The size of the horizontal detector cannot be negative or zero")
        if detectors_x_pad < 0:
            raise ValueError("The padding size of the horizontal detector cannot be negative")
        angles_vec = np.asarray(angles_vec)
        if angles_vec.size == 0:
            raise ValueError("The length of angles array cannot be zero")
        if angles_vec.ndim >= 2:
            raise ValueError("The array of angles must be 1D")
        if centre_of_rotation is None:
            centre_of_rotation = 0.0
        if np.ndim(centre_of_rotation) == 1 and len(centre_of_rotation) != len(angles_vec):
            raise ValueError("The CoR must be a scalar or a 1D array of the SAME size as angles")
        if np.ndim(centre_of_rotation) > 1:
            raise ValueError("A CoR with a vertical component breaks slice independence and is not supported")
        if isinstance(recon_size, tuple):
            raise ValueError(
                "Reconstruction is currently available for squared or cubic objects only, please provide a scalar"
            )
        if recon_size <= 0:
            raise ValueError("The size of the reconstruction object cannot be zero")
        if processing_arch != "gpu":
            raise ValueError("3D CPU reconstruction is not supported, please use GPU")
        if device_index is None or device_index < 0:
            raise ValueError("The GPU device index must be zero or positive (there is no CPU path)")
        if ordsub_number is None:
            ordsub_number = 1
        if ordsub_number <= 0:
            raise ValueError("The number of ordered subsets cannot be negative or zero")
        if detectors_y is None or detectors_y <= 0:
            raise ValueError("The size of the vertical detector cannot be negative or zero")

        self.detectors_x = int(detectors_x)
        self.detectors_x_pad = int(detectors_x_pad)
        self.detectors_y = int(detectors_y)
        self.angles_vec = angles_vec
        self.centre_of_rotation = centre_of_rotation
        self.recon_size = int(recon_size)
        self.processing_arch = processing_arch
        self.device_index = int(device_index)
        self.ordsub_number = int(ordsub_number)
        self.fbp_filter_type = "ram-lak"
        self.fbp_filter_parameter = None
        self.fbp_filter_d = None
        self.device = torch.device("cuda", self.device_index)

        self.nu = self.detectors_x + 2 * self.detectors_x_pad
        na = int(angles_vec.size)
        # geom_size equivalents (astra.geom_size(vol_geom) / (proj_geom))
        self.vol_geom = (self.detectors_y, self.recon_size, self.recon_size)
        self.proj_geom = (self.detectors_y, na, self.nu)

        # sin/cos in the dtype of AnglesVec, like np.cos(theta) in supp/funcs.py:74-81
        cos_t = np.ascontiguousarray(np.cos(angles_vec), dtype=np.float64)
        sin_t = np.ascontiguousarray(np.sin(angles_vec), dtype=np.float64)
        cor = np.ascontiguousarray(np.broadcast_to(np.asarray(centre_of_rotation, dtype=np.float64), (na,)))
        dp = C.POINTER(C.c_double)
        self._g = lib.tmb_geom_create(
            self.detectors_y, self.recon_size, self.nu, na,
            cos_t.ctypes.data_as(dp), sin_t.ctypes.data_as(dp), cor.ctypes.data_as(dp),
            self.ordsub_number, 1 if quantise_weights else 0,
        )
        if not self._g:
            raise ValueError(lib.tmb_last_error().decode())

        # ordered-subset table in the reference's format (astra_base.py:195-209)
        if self.ordsub_number > 1:
            self.NumbProjBins = int(np.ceil(float(na) / float(self.ordsub_number)))
            self.newInd_Vec = np.zeros([self.ordsub_number, self.NumbProjBins], dtype="int")
            row = (C.c_int * self.NumbProjBins)()
            for s in range(self.ordsub_number):
                check(lib.tmb_geom_subset_row(self._g, s, row), "tmb_geom_subset_row")
                self.newInd_Vec[s, :] = np.frombuffer(row, dtype=np.int32)
            self.proj_geom_OS = {
                s: (self.detectors_y, lib.tmb_geom_subset_size(self._g, s), self.nu)
                for s in range(self.ordsub_number)
            }
        self._ws = None
        if verbosity:
            print("3D <gpu> parallel-beam projection geometry initialised (libtmb)...")

    def __del__(self):
        g = getattr(self, "_g", None)
        if g:
            lib.tmb_geom_destroy(g)
            self._g = None

    # ---- scratch ---------------------------------------------------------------------------
    def _workspace(self) -> torch.Tensor:
        if self._ws is None:
            nbytes = lib.tmb_geom_workspace_bytes(self._g)
            # zero-filled once: the kernels keep the zero borders of the interior layouts intact
            self._ws = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        return self._ws

    def free_workspace(self) -> None:
        self._ws = None

    def angle_table(self) -> np.ndarray:
        out = np.empty((self.angles_vec.size, 8), dtype=np.float32)
        check(lib.tmb_geom_table(self._g, out.ctypes.data_as(C.POINTER(C.c_float))), "tmb_geom_table")
        return out

    def subset_size(self, os_index: Optional[int]) -> int:
        return lib.tmb_geom_subset_size(self._g, -1 if os_index is None else os_index)

    def _sub(self, os_index: Optional[int]) -> int:
        if os_index is None or self.ordsub_number == 1:
            return -1
        if not 0 <= os_index < self.ordsub_number:
            raise ValueError(f"subset index {os_index} out of range")
        return int(os_index)

    # ---- A and A^T ---------------------------------------------------------------------------
    def _forward(self, vol, sub: int) -> torch.Tensor:
        vol = require_dense(as_cuda_f32(vol, self.device, "volume"), self.vol_geom, "volume")
        na_s = lib.tmb_geom_subset_size(self._g, sub)
        out = torch.empty((self.detectors_y, na_s, self.nu), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.tmb_fp3d(self._g, sub, ptr(vol), ptr(out), ptr(self._workspace()), stream_ptr(out)),
                  "tmb_fp3d")
        return out

    def _backward(self, sino, sub: int) -> torch.Tensor:
        na_s = lib.tmb_geom_subset_size(self._g, sub)
        sino = require_dense(as_cuda_f32(sino, self.device, "projection data"),
                             (self.detectors_y, na_s, self.nu), "projection data")
        out = torch.empty(self.vol_geom, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.tmb_bp3d(self._g, sub, ptr(sino), ptr(out), ptr(self._workspace()), stream_ptr(out)),
                  "tmb_bp3d")
        return out

    def _forwprojCuPy(self, object3D) -> torch.Tensor:
        """astra_tools3d.py:78-81"""
        return self._forward(object3D, -1)

    def _forwprojOSCuPy(self, object3D, os_index: int) -> torch.Tensor:
        """astra_tools3d.py:83-86"""
        return self._forward(object3D, self._sub(os_index))

    def _backprojCuPy(self, proj_data) -> torch.Tensor:
        """astra_tools3d.py:102-105"""
        return self._backward(proj_data, -1)

    def _backprojOSCuPy(self, proj_data, os_index: int) -> torch.Tensor:
        """astra_tools3d.py:107-110"""
        return self._backward(proj_data, self._sub(os_index))

    # ---- fused gradient of the data term (data_fidelities.py:7-40) ---------------------------
    def grad_data_term(self, x: torch.Tensor, b_full: torch.Tensor, os_index: Optional[int],
                       fidelity: str = "LS", w_full: Optional[torch.Tensor] = None,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
        sub = self._sub(os_index)
        x = require_dense(as_cuda_f32(x, self.device, "volume"), self.vol_geom, "volume")
        b_full = require_dense(b_full, self.proj_geom, "projection data")
        if w_full is not None:
            w_full = require_dense(w_full, self.proj_geom, "weights")
        if out is None:
            out = torch.empty(self.vol_geom, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.tmb_grad(self._g, sub, FIDELITY[fidelity], ptr(x), ptr(b_full),
                               ptr(w_full) if w_full is not None else None, ptr(out),
                               ptr(self._workspace()), stream_ptr(out)), "tmb_grad")
        return out

    # ---- robust / ring-artefact data terms (extension; include/tmb.h tmb_grad_ext) ----------------
    def grad_data_term_ext(self, x: torch.Tensor, b_full: torch.Tensor, os_index: Optional[int],
                           fidelity: str = "LS", w_full: Optional[torch.Tensor] = None,
                           huber_threshold: Optional[float] = None, ring_rx: Optional[torch.Tensor] = None,
                           ring_alpha: float = 0.0, beta_swls: float = 0.0,
                           ring_vec: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                           studentst_threshold: Optional[float] = None) -> torch.Tensor:
        """grad = A_s^T rho'(A_s x - b_s) for the Huber / Student's-t / Group-Huber ring / SWLS models; ``ring_vec``
        ([nz, nu]) receives the angle-sum of the ring-corrected residual."""
        sub = self._sub(os_index)
        mode = {"LS": 0, "PWLS": 1, "SWLS": 2}[fidelity]
        x = require_dense(as_cuda_f32(x, self.device, "volume"), self.vol_geom, "volume")
        b_full = require_dense(b_full, self.proj_geom, "projection data")
        if mode:
            w_full = require_dense(w_full, self.proj_geom, "weights")
        if ring_rx is not None:
            ring_rx = require_dense(ring_rx, (self.detectors_y, self.nu), "ring variable")
            ring_vec = require_dense(ring_vec, (self.detectors_y, self.nu), "ring residual sum")
        if out is None:
            out = torch.empty(self.vol_geom, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.tmb_grad_ext(self._g, sub, ptr(x), ptr(b_full), ptr(w_full) if mode else None, mode,
                                   float(huber_threshold or 0.0), float(studentst_threshold or 0.0),
                                   ptr(ring_rx) if ring_rx is not None else None,
                                   float(ring_alpha), float(beta_swls),
                                   ptr(ring_vec) if ring_rx is not None else None, ptr(out),
                                   ptr(self._workspace()), stream_ptr(out)), "tmb_grad_ext")
        return out
