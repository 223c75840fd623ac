"""Host-array (numpy in, numpy out) entry points of the direct methods: the 3-D part of the reference's
``RecToolsDIR`` (tomobar/methodsDIR.py:18-175) over the same CUDA path as ``RecToolsDIRCuPy``.

The reference class drives ASTRA with host arrays (its 3-D FBP is ``_backproj(_filtersinc3D(data))`` with
the sinc parameter a = 1.1, methodsDIR.py:171-175, :257-292); its pinned results are those of the CuPy class
with ``cutoff_freq=1.1`` (tests/test_RecToolsDIR.py:265-323 vs tests/test_RecToolsDIRCuPy.py:543-566).  Here
the arrays are staged to the GPU, run through ``RecToolsDIRCuPy`` and copied back.

2-D data (``DetectorsDimV`` 0 or None; ``AstraTools2D``, astra_wrappers/astra_tools2d.py:8): a single slice through
the same kernels.  ASTRA's 2-D image is y-up -- row 0 is the TOP row -- i.e. the vertical flip of a slice of the 3-D
path ("orientation"; pinned on the reference's CPU golden by the config-1 tests), so
images are flipped on the way in and out.  2-D ``FBP`` follows the reference's GPU branch (ASTRA's FBP_CUDA,
methodsDIR.py:150-158): ramp filter on the detector zero-padded to a power of two, windows ``ram-lak``,
``shepp-logan``, ``cosine``, ``hamming``, ``hann`` with the cut-off ``filter_d``, scaled by pi / (2 angles); the
reference pins only value RANGES for it (tests/test_RecToolsDIR.py:65-169), so its parity is unpinned.

Not provided: the ``device_projector="cpu"`` arch (there is no CPU path in this package; the reference's CPU FBP is
restated by the test infrastructure as the config-1 baseline) and the scipy ``FOURIER`` method (2-D CPU gridding, not part of the
GPU hot path).
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
        self.geom = "2D" if (DetectorsDimV == 0 or DetectorsDimV is None) else "3D"
        self._gpu = RecToolsDIRCuPy(DetectorsDimH, DetectorsDimH_pad, 1 if self.geom == "2D" else DetectorsDimV,
                                    CenterRotOffset, AnglesVec, ObjSize, projector=projector, device_projector=gpu_index)
        self.Atools = self._gpu.Atools
        # ASTRA 2-D FBP specific parameters (astra_base.py:70-72)
        self.Atools.fbp_filter_type, self.Atools.fbp_filter_parameter, self.Atools.fbp_filter_d = "ram-lak", None, None

    @staticmethod
    def _host(data, what: str) -> np.ndarray:
        data = np.asarray(data)
        if data.dtype != np.float32:
            raise ValueError(f"The {what} should be float32 data type")
        return data

    def FORWPROJ(self, data: np.ndarray, **kwargs) -> np.ndarray:
        """Forward projection of a 3-D object; ``data_axes_labels_order`` orders the OUTPUT
        (methodsDIR.py:71-94)."""
        if self.geom == "2D":
            image = self._host(data, "object")
            if image.ndim != 2:
                raise ValueError("RecToolsDIR.FORWPROJ: a 2-D image expected for the 2-D geometry")
            sino = self._gpu.FORWPROJ(np.ascontiguousarray(image[::-1])[None])[0]  # y-up image -> slice of the 3-D path
            labels = kwargs.get("data_axes_labels_order")
            if labels is not None:
                sino = _data_dims_swapper(sino, labels, ["angles", "detX"])
            return np.ascontiguousarray(sino.cpu().numpy())
        out = self._gpu.FORWPROJ(np.ascontiguousarray(self._host(data, "object")), **kwargs)
        return np.ascontiguousarray(out.cpu().numpy())

    def _sino2d(self, data, kwargs) -> np.ndarray:
        sino = self._host(data, "projection data")
        if sino.ndim != 2:
            raise ValueError("RecToolsDIR: 2-D projection data [angles, detX] expected for the 2-D geometry")
        labels = kwargs.get("data_axes_labels_order")
        if labels is not None:
            sino = np.transpose(sino, [list(labels).index(a) for a in ("angles", "detX")])
        return np.ascontiguousarray(sino)

    def BACKPROJ(self, data: np.ndarray, **kwargs) -> np.ndarray:
        """Back-projection of 3-D projection data, default axes ["detY", "angles", "detX"]
        (methodsDIR.py:96-119)."""
        if self.geom == "2D":
            rec = self._gpu.BACKPROJ(self._sino2d(data, kwargs)[None])[0]
            return np.ascontiguousarray(rec.cpu().numpy()[::-1])
        return self._gpu.BACKPROJ(self._host(data, "projection data"), **kwargs).cpu().numpy()

    def FBP(self, data: np.ndarray, **kwargs) -> np.ndarray:
        """3-D filtered back-projection with the customised sinc filter, a = 1.1 (methodsDIR.py:121-175).
        Default axes ["detY", "angles", "detX"]; ``recon_mask_radius`` as in ``check_kwargs``.  The
        ``filter_type`` / ``filter_parameter`` / ``filter_d`` keywords only reach ASTRA's 2-D FBP_CUDA in the
        reference and are accepted and ignored here."""
        if self.geom == "2D":
            return self._fbp2d(data, kwargs)
        data = self._host(data, "projection data")
        labels = kwargs.get("data_axes_labels_order")
        if labels is None:
            labels = _ACCEPTED
        if data.ndim != 3:
            raise ValueError("RecToolsDIR.FBP: 3-D projection data expected")
        passed = {k: v for k, v in kwargs.items() if k in ("recon_mask_radius",)}
        rec = self._gpu.FBP(data, data_axes_labels_order=list(labels), cutoff_freq=1.1, **passed)
        return rec.cpu().numpy()

    _WINDOWS = ("ram-lak", "shepp-logan", "cosine", "hamming", "hann")

    def _fbp2d(self, data, kwargs) -> np.ndarray:
        """The GPU branch of the reference's 2-D FBP (methodsDIR.py:150-158 -> ASTRA FBP_CUDA, astra_base.py:311-370)."""
        import torch

        from tomobar_b200.supp.suppTools import _apply_horiz_detector_padding, check_kwargs

        A = self.Atools
        for key in ("filter_type", "filter_parameter", "filter_d"):
            if key in kwargs:
                setattr(A, "fbp_" + key, kwargs[key])
        sino = torch.from_numpy(self._sino2d(data, kwargs)).to(A.device)
        sino = _apply_horiz_detector_padding(sino[None], A.detectors_x_pad, True)[0]
        na, nu = sino.shape
        width = 2 ** int(np.ceil(np.log2(2 * nu)))
        freq = torch.fft.rfftfreq(width, device=sino.device).to(torch.float32) * 2.0  # 0 .. 1 (Nyquist)
        window = A.fbp_filter_type if A.fbp_filter_type in self._WINDOWS else "ram-lak"
        if window != A.fbp_filter_type:
            print(f"FBP: filter {A.fbp_filter_type!r} is not built, ram-lak is used")
        d = 1.0 if A.fbp_filter_d is None else float(A.fbp_filter_d)
        arg = freq / d
        filt = freq.clone()
        if window == "shepp-logan":
            filt *= torch.sinc(arg / 2.0)
        elif window == "cosine":
            filt *= torch.cos(arg * (np.pi / 2.0))
        elif window == "hamming":
            filt *= 0.54 + 0.46 * torch.cos(arg * np.pi)
        elif window == "hann":
            filt *= 0.5 + 0.5 * torch.cos(arg * np.pi)
        if window != "ram-lak":
            filt = torch.where(freq > d, torch.zeros_like(filt), filt)
        filtered = torch.fft.irfft(torch.fft.rfft(sino, n=width, dim=1) * filt, n=width, dim=1)[:, :nu]
        rec = A._backprojCuPy(filtered.contiguous()[None]) * np.float32(np.pi / (2.0 * na))
        rec = torch.flip(rec, dims=(1,))  # back to ASTRA's y-up 2-D image
        passed = {k: v for k, v in kwargs.items() if k == "recon_mask_radius"}
        rec = check_kwargs(rec, cupyrun=True, **passed)
        return np.ascontiguousarray(rec[0].cpu().numpy())

    def FOURIER(self, data: np.ndarray, **kwargs) -> np.ndarray:
        raise NotImplementedError("RecToolsDIR.FOURIER (2-D scipy gridding on the CPU) is outside the GPU hot path; "
                                  "use RecToolsDIRCuPy.FOURIER_INV")


__all__ = ["RecToolsDIR", "_data_dims_swapper"]
