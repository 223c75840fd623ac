"""Direct reconstruction on B200 behind the ``RecToolsDIRCuPy`` interface
(tomobar/methodsDIR_CuPy.py:26-150): FORWPROJ, BACKPROJ, FBP.  Arrays are float32 CUDA torch
tensors.  FOURIER_INV is provided by ``tomobar_b200.fourier_inv`` when built."""

from __future__ import annotations

from typing import Literal

import numpy as np
import torch

from tomobar_b200._tensors import as_cuda_f32
from tomobar_b200.fourier import _filtersinc3D_cupy
from tomobar_b200.projector import ProjTools3D
from tomobar_b200.supp.funcs import _data_dims_swapper
from tomobar_b200.supp.suppTools import _apply_horiz_detector_padding, check_kwargs


class RecToolsDIRCuPy:
    """Direct methods (methodsDIR_CuPy.py:39-68).

    Args:
        DetectorsDimH (int): Horizontal detector dimension.
        DetectorsDimH_pad (int): The amount of padding for the horizontal detector.
        DetectorsDimV (int): Vertical detector dimension for 3D case, 0 or None for 2D case.
        CenterRotOffset (float, ndarray): Centre of Rotation scalar or one value per angle.
        AnglesVec (np.ndarray): Vector of projection angles in radians.
        ObjSize (int): Reconstructed object dimensions (a scalar).
        projector: kept for signature compatibility ("astra" | "fourier"); both run in libtmb.
        device_projector (int): GPU index.
    """

    def __init__(
        self,
        DetectorsDimH,
        DetectorsDimH_pad,
        DetectorsDimV,
        CenterRotOffset,
        AnglesVec,
        ObjSize,
        projector: Literal["fourier", "astra"] = "astra",
        device_projector=0,
        quantise_weights: bool = True,
    ):
        self.detectors_x_pad = DetectorsDimH_pad
        if CenterRotOffset is None:
            CenterRotOffset = 0.0
        self.centre_of_rotation = CenterRotOffset
        self.angles_vec = AnglesVec
        self.recon_size = ObjSize
        self.projector = projector
        if DetectorsDimV == 0 or DetectorsDimV is None:
            DetectorsDimV = 1
        if DetectorsDimH_pad > 0:
            # padded detector => padded reconstruction grid, like methodsDIR.py's parent class
            obj = ObjSize
        else:
            obj = ObjSize
        self.Atools = ProjTools3D(
            DetectorsDimH,
            DetectorsDimH_pad,
            DetectorsDimV,
            AnglesVec,
            CenterRotOffset,
            obj,
            "gpu",
            device_projector if isinstance(device_projector, int) else 0,
            None,
            quantise_weights=quantise_weights,
        )

    def FORWPROJ(self, data, **kwargs) -> torch.Tensor:
        """Forward projection of a volume [detY, N, N] (methodsDIR_CuPy.py:70-88)."""
        projected = self.Atools._forwprojCuPy(data)
        for key, value in kwargs.items():
            if key == "data_axes_labels_order" and value is not None:
                projected = _data_dims_swapper(projected, value, ["detY", "angles", "detX"])
        return projected

    def BACKPROJ(self, data, **kwargs) -> torch.Tensor:
        """Back-projection of projection data (methodsDIR_CuPy.py:90-112).  The logical array is
        back-projected; the reference hands ASTRA the raw pointer of the swapped view."""
        data = as_cuda_f32(data, self.Atools.device, "projection data")
        for key, value in kwargs.items():
            if key == "data_axes_labels_order" and value is not None:
                data = _data_dims_swapper(data, value, ["detY", "angles", "detX"])
        data = _apply_horiz_detector_padding(data, self.Atools.detectors_x_pad, True)
        return self.Atools._backprojCuPy(data)

    def FBP(self, data, **kwargs) -> torch.Tensor:
        """Filtered back-projection with the sinc filter (methodsDIR_CuPy.py:114-150).
        Input axes default to ["angles", "detY", "detX"]."""
        kwargs.update({"cupyrun": True})
        cutoff_freq = 0.35
        data = as_cuda_f32(data, self.Atools.device, "projection data")
        for key, value in kwargs.items():
            if key == "data_axes_labels_order" and value is not None:
                data = _data_dims_swapper(data, value, ["angles", "detY", "detX"])
            if key == "cutoff_freq" and value is not None:
                cutoff_freq = value
        data = _apply_horiz_detector_padding(data, self.Atools.detectors_x_pad, True)
        data = _filtersinc3D_cupy(data.contiguous(), cutoff=cutoff_freq)
        data = data.swapaxes(0, 1).contiguous()
        reconstruction = self.Atools._backprojCuPy(data)
        return check_kwargs(reconstruction, **kwargs)
