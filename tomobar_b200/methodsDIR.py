"""Host-array (numpy in, numpy out) entry points of the direct methods: the 3-D part of the reference's
``RecToolsDIR`` (tomobar/methodsDIR.py:18-175) over the same CUDA path as ``RecToolsDIRCuPy``.

The reference class drives ASTRA with host arrays (its 3-D FBP is ``_backproj(_filtersinc3D(data))`` with
the sinc parameter a = 1.1, methodsDIR.py:171-175, :257-292); its pinned results are those of the CuPy class
with ``cutoff_freq=1.1`` (tests/test_RecToolsDIR.py:265-323 vs tests/test_RecToolsDIRCuPy.py:543-566).  Here
the arrays are staged to the GPU, run through ``RecToolsDIRCuPy`` and copied back.  Not provided: the
``device_projector="cpu"`` arch (there is no CPU path in this package), the 2-D geometry (ASTRA's y-up 2-D
class) and the scipy ``FOURIER`` method (2-D CPU gridding, not part of the GPU hot path).
"""

from __future__ import annotations

from typing import Literal

import numpy as np

from tomobar_b200.methodsDIR_CuPy import RecToolsDIRCuPy
from tomobar_b200.supp.funcs import _data_dims_swapper, _parse_device_argument

_ACCEPTED = ["detY", "angles", "detX"]


class RecToolsDIR:
    """Direct reconstruction from host arrays (methodsDIR.py:18-69).

    Args:
        DetectorsDimH (int): Horizontal detector dimension.
        DetectorsDimH_pad (int): The amount of padding for the horizontal detector.
        DetectorsDimV (int): Vertical detector dimension (3-D only).
        CenterRotOffset (float, ndarray): Centre of Rotation scalar or one value per angle.
        AnglesVec (np.ndarray): Vector of projection angles in radians.
        ObjSize (int): Reconstructed object dimensions (a scalar).
        projector: kept for signature compatibility.
        device_projector: "gpu" or a GPU index.
    """

    def __init__(self, DetectorsDimH, DetectorsDimH_pad, DetectorsDimV, CenterRotOffset, AnglesVec, ObjSize,
                 projector: Literal["fourier", "astra"] = "astra", device_projector="gpu"):
        arch, gpu_index = _parse_device_argument(device_projector)
        if arch != "gpu":
            raise ValueError('tomobar_b200 has no CPU projector: use device_projector="gpu" or a GPU index')
        if DetectorsDimV == 0 or DetectorsDimV is None:
            raise NotImplementedError("RecToolsDIR: the 2-D geometry class is not built (3-D host arrays only)")
        self.geom = "3D"
        self._gpu = RecToolsDIRCuPy(DetectorsDimH, DetectorsDimH_pad, DetectorsDimV, CenterRotOffset, AnglesVec, ObjSize,
                                    projector=projector, device_projector=gpu_index)
        self.Atools = self._gpu.Atools

    @staticmethod
    def _host(data, what: str) -> np.ndarray:
        data = np.asarray(data)
        if data.dtype != np.float32:
            raise ValueError(f"The {what} should be float32 data type")
        return data

    def FORWPROJ(self, data: np.ndarray, **kwargs) -> np.ndarray:
        """Forward projection of a 3-D object; ``data_axes_labels_order`` orders the OUTPUT
        (methodsDIR.py:71-94)."""
        out = self._gpu.FORWPROJ(np.ascontiguousarray(self._host(data, "object")), **kwargs)
        return np.ascontiguousarray(out.cpu().numpy())

    def BACKPROJ(self, data: np.ndarray, **kwargs) -> np.ndarray:
        """Back-projection of 3-D projection data, default axes ["detY", "angles", "detX"]
        (methodsDIR.py:96-119)."""
        return self._gpu.BACKPROJ(self._host(data, "projection data"), **kwargs).cpu().numpy()

    def FBP(self, data: np.ndarray, **kwargs) -> np.ndarray:
        """3-D filtered back-projection with the customised sinc filter, a = 1.1 (methodsDIR.py:121-175).
        Default axes ["detY", "angles", "detX"]; ``recon_mask_radius`` as in ``check_kwargs``.  The
        ``filter_type`` / ``filter_parameter`` / ``filter_d`` keywords only reach ASTRA's 2-D FBP_CUDA in the
        reference and are accepted and ignored here."""
        data = self._host(data, "projection data")
        labels = kwargs.get("data_axes_labels_order")
        if labels is None:
            labels = _ACCEPTED
        if data.ndim != 3:
            raise ValueError("RecToolsDIR.FBP: 3-D projection data expected")
        passed = {k: v for k, v in kwargs.items() if k in ("recon_mask_radius",)}
        rec = self._gpu.FBP(data, data_axes_labels_order=list(labels), cutoff_freq=1.1, **passed)
        return rec.cpu().numpy()

    def FOURIER(self, data: np.ndarray, **kwargs) -> np.ndarray:
        raise NotImplementedError("RecToolsDIR.FOURIER (2-D scipy gridding on the CPU) is outside the GPU hot path; "
                                  "use RecToolsDIRCuPy.FOURIER_INV")


__all__ = ["RecToolsDIR", "_data_dims_swapper"]
