"""Direct reconstruction on B200 behind the ``RecToolsDIRCuPy`` interface
(tomobar/methodsDIR_CuPy.py:26-447): FORWPROJ, BACKPROJ, FBP, FOURIER_INV.  Arrays are float32
CUDA torch tensors."""

from __future__ import annotations

import math
from typing import Literal

import numpy as np
import torch

from tomobar_b200._lib import lib, check
from tomobar_b200._tensors import as_cuda_f32, ptr, stream_ptr
from tomobar_b200.fourier import _filtersinc3D_cupy, calc_filter
from tomobar_b200.projector import ProjTools3D
from tomobar_b200.supp.funcs import _raw_buffer_view, _data_dims_swapper, _parse_device_argument
from tomobar_b200.supp.suppTools import _apply_horiz_detector_padding, check_kwargs, edge_pad


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
        # True: BACKPROJ reproduces the reference's non-contiguous-view behaviour
        self.compat_view_bug = False
        if CenterRotOffset is None:
            CenterRotOffset = 0.0
        self.centre_of_rotation = CenterRotOffset
        self.angles_vec = AnglesVec
        self.recon_size = ObjSize
        self.projector = projector
        if DetectorsDimV == 0 or DetectorsDimV is None:
            DetectorsDimV = 1
        arch, gpu_index = _parse_device_argument(device_projector)
        self.Atools = ProjTools3D(
            DetectorsDimH,
            DetectorsDimH_pad,
            DetectorsDimV,
            AnglesVec,
            CenterRotOffset,
            ObjSize,
            arch,
            gpu_index,
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
        if self.compat_view_bug and self.Atools.detectors_x_pad == 0:
            # bug-for-bug with the reference (its golden tests/test_RecToolsDIRCuPy.py:714-715 encodes it): the
            # swapped VIEW's raw buffer is back-projected with the view's logical shape
            data = _raw_buffer_view(data)
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

    # ------------------------------------------------------------------------------------------
    _FILTERS = ("none", "ramp", "shepp", "cosine", "cosine2", "hamming", "hann", "parzen")
    _CENTER_SIZE_MIN = 192  # methodsDIR_CuPy.py:23
    _GATHER_SLICE_PAIRS = True  # FOURIER_INV's whole-grid gather reads polar samples stored as slice pairs (False: planar)
    _FILTER_SLICE_PAIRS = True  # FOURIER_INV filters slice pairs as complex rows (False: an r2c / c2r pair per slice)

    def FOURIER_INV(self, data, **kwargs) -> torch.Tensor:
        """Direct Fourier inversion on unequally spaced grids (USFFT gridding, Nikitin's
        "fourierrec"; methodsDIR_CuPy.py:152-447, default centre-gather path).

        Keyword Args:
            data_axes_labels_order (list, None): axes of the input; default ["detY", "angles", "detX"].
            recon_mask_radius (float): circular mask radius.
            filter_type (str): none, ramp, shepp, cosine, cosine2, hamming, hann, parzen.
            cutoff_freq (float): filter cutoff (default 1.0).
            padding (int): extra zero padding of the frequency grid.
            power_of_2_oversampling / power_of_2_cropping (bool): as in the reference.
            center_size (int): side of the centre square of the frequency grid that is gathered; the rest is
                scattered with atomic adds, everything when it is below 192 (default: the whole grid is gathered).
        The memory-tuning keywords of the reference (chunk_count, min_mem_usage_*, block_dim*) are
        accepted and ignored.
        """
        if isinstance(data, tuple):
            # dry run under an active DeviceMemStack: record this implementation's allocations
            # (the reference's estimator mode, methodsDIR_CuPy.py:253-258, 437-441)
            return self._fourier_inv_estimator(data, **kwargs)
        kwargs.update({"cupyrun": True})
        cutoff_freq = 1.0
        filter_type = "shepp"
        oversampling_level = 4
        power_of_2_oversampling = True
        power_of_2_cropping = False
        padding = 0
        center_size = 32768
        data = as_cuda_f32(data, self.Atools.device, "projection data")
        for key, value in kwargs.items():
            if value is None:
                continue
            if key == "data_axes_labels_order":
                data = _data_dims_swapper(data, value, ["detY", "angles", "detX"])
            elif key == "center_size":
                center_size = int(value)
            elif key == "cutoff_freq":
                cutoff_freq = value
            elif key == "filter_type":
                if value not in self._FILTERS:
                    print("Unknown filter name, please use: none, ramp, shepp, cosine, cosine2, hamming, hann or "
                          "parzen. Set to shepp filter")
                else:
                    filter_type = value
            elif key == "power_of_2_oversampling":
                power_of_2_oversampling = value
            elif key == "power_of_2_cropping":
                power_of_2_cropping = value
            elif key == "padding":
                if not isinstance(value, int) or value < 0:
                    print(f"Invalid padding: {value}. Set to 0")
                else:
                    padding = value

        dev = data.device
        nz, nproj, data_n = data.shape
        recon_size = self.recon_size
        if recon_size > data_n:
            raise ValueError(
                "The reconstruction size {} should not be larger than the size of the horizontal detector {}".format(
                    recon_size, data_n
                )
            )
        # odd sizes are padded to even: slices are processed in pairs (:268-282)
        odd_horiz, odd_vert = bool(data_n % 2), bool(nz % 2)
        data_n += odd_horiz
        nz += odd_vert
        if odd_horiz or odd_vert:
            data_p = torch.zeros((nz, nproj, data_n), dtype=torch.float32, device=dev)
            data_p[: nz - odd_vert, :, : data_n - odd_horiz] = data
            if odd_horiz:
                data_p[: nz - odd_vert, :, -1] = data[..., -1]
            data = data_p
        data = data.contiguous()

        n = data_n + self.detectors_x_pad * 2 + padding * 2
        if power_of_2_cropping:
            n_pow2 = 2 ** math.ceil(math.log2(n))
            if 0.9 < n / n_pow2:
                n = n_pow2
        nz2 = nz // 2

        theta = torch.as_tensor(-np.asarray(self.angles_vec), dtype=torch.float32, device=dev)
        sorted_theta, sorted_idx = torch.sort(theta)
        sorted_idx = sorted_idx.to(torch.int32)

        eps = 1e-4  # accuracy of the USFFT
        mu = -np.log(eps) / (2 * n * n)
        st = torch.cuda.current_stream(dev).cuda_stream

        with torch.cuda.device(dev):
            # STEP 0: filtering on an oversampled detector with a half-pixel phase ramp (:449-545)
            # ... and STEP 1a: slices paired into complex slices (:645-683), packed chunk by chunk out of the filter output
            datac = torch.empty((nz2, nproj, n), dtype=torch.complex64, device=dev)
            self._fourier_filter(data, data_n, n, power_of_2_oversampling, oversampling_level, filter_type, cutoff_freq,
                                 pack_into=datac)
            del data
            # STEP 1b: 1-D FFT along the detector (:725-754)
            datac = torch.fft.fft(datac, dim=-1)
            center_size = min(center_size, 2 * n)
            center_size -= center_size % 2
            whole = center_size >= self._CENTER_SIZE_MIN and center_size == 2 * n
            partial = center_size >= self._CENTER_SIZE_MIN
            chunk = max(1, min(nz2, (1 << 28) // (4 * n * n)))  # complex slices per pass of STEPS 2-4
            # the whole-grid gather reads slice PAIRS (one 128-bit load per two slices) when every chunk is whole
            # blocks of 8 complex slices: the scale / sign pass writes that layout instead of working in place
            pairs = whole and self._GATHER_SLICE_PAIRS and nz2 % 8 == 0 and chunk % 8 == 0
            if pairs:
                dataz = torch.empty_like(datac)
                check(lib.tmb_fi_scale_sign_pairs(ptr(datac), ptr(dataz), float(np.float32(4 / n)), n, nproj, nz2, st),
                      "tmb_fi_scale_sign_pairs")
                datac = dataz
                del dataz
            else:
                check(lib.tmb_fi_scale_sign(ptr(datac), float(np.float32(4 / n)), n, nproj, nz2, st), "tmb_fi_scale_sign")
            m = int(np.ceil(2 * n * 1 / np.pi * np.sqrt(-mu * np.log(eps) + (mu * n) * (mu * n) / 4)))
            # STEP 2: polar samples onto the 2n x 2n Cartesian grid (:756-835); the (-1)^(x+y) before the 2-D
            # FFT is applied by these kernels, the one after it by the unpadding kernel.  Three branches, chosen
            # by the centre size like the reference: the whole grid gathered (default), a centre square gathered
            # and the rest scattered with atomic adds, or everything scattered (centre below _CENTER_SIZE_MIN)
            # STEP 4's geometry: crop, de-apodise, unpack the slice pairs (:920-966)
            odd_recon = bool(recon_size % 2)
            unpad_z = nz - odd_vert
            um = (n - odd_horiz) // 2 - recon_size // 2
            up = (n - odd_horiz) // 2 + (recon_size + odd_recon) // 2
            rs = up - um
            recon_up = torch.empty((unpad_z, rs, rs), dtype=torch.float32, device=dev)
            # STEPS 2-4 run chunk by chunk of complex slices (the slices are independent): the oversampled grid only
            # ever exists for one chunk, the inverse 2-D FFT's output is read directly by the unpadding kernel (no
            # copy back into a whole-volume grid) and its 1 / (2n)^2 is applied there (no normalisation pass)
            mu32, inv_grid = float(np.float32(mu)), float(np.float32(1.0 / (4.0 * n * n)))
            for s0 in range(0, nz2, chunk):
                c = min(chunk, nz2 - s0)
                dc = ptr(datac[s0:])
                if whole:
                    fde = torch.empty((c, 2 * n, 2 * n), dtype=torch.complex64, device=dev)
                    gather = lib.tmb_fi_gather_pairs if pairs else lib.tmb_fi_gather
                    check(gather(dc, ptr(fde), ptr(theta), ptr(sorted_theta), ptr(sorted_idx), m, mu32, n, nproj, c, st),
                          "tmb_fi_gather")
                else:
                    # (the reference adds onto cp.empty memory in the partial branch, :661-670; zeros are what it means)
                    fde = torch.zeros((c, 2 * n, 2 * n), dtype=torch.complex64, device=dev)
                    check(lib.tmb_fi_scatter(dc, ptr(fde), ptr(theta), m, mu32, center_size if partial else 0, n, nproj,
                                             c, st), "tmb_fi_scatter")
                    if partial:
                        check(lib.tmb_fi_gather_center(dc, ptr(fde), ptr(theta), ptr(sorted_theta), ptr(sorted_idx), m,
                                                       mu32, n, nproj, c, center_size, st), "tmb_fi_gather_center")
                # STEP 3: centred 2-D inverse FFT (:851-896), unnormalised
                fde = torch.fft.ifft2(fde, dim=(-2, -1), norm="forward")
                check(lib.tmb_fi_unpad(ptr(recon_up[2 * s0:]), ptr(fde), mu32, inv_grid, nproj, up, unpad_z - 2 * s0, um, n,
                                       c, st), "tmb_fi_unpad")
                del fde
            del datac
        return check_kwargs(recon_up, **kwargs)

    def _fourier_inv_estimator(self, shape, **kwargs):
        """FOURIER_INV called with a shape tuple (after any axis swap: [detY, angles, detX]): replays the
        allocation sequence of the method above on the active ``DeviceMemStack`` and returns the shape of
        the reconstruction.  The stages are the reference's ``*_estimator`` twins (methodsDIR_CuPy.py:547, 685,
        837, 898, 968), each replaying THIS implementation's allocations of that stage.  cuFFT work areas are
        counted as one copy of the transform's output (an upper bound for the power-of-two sizes used here); the
        caller's input array is counted like the reference counts it (``data_dtype`` keyword, default float32)."""
        from tomobar_b200.supp.memory_estimator_helpers import DeviceMemStack

        stack = DeviceMemStack.instance()
        if stack is None:
            raise ValueError("FOURIER_INV: a shape tuple needs an active DeviceMemStack (memory estimation)")
        power_of_2_oversampling, power_of_2_cropping, oversampling_level, padding = True, False, 4, 0
        for key, value in kwargs.items():
            if value is None:
                continue
            if key == "data_axes_labels_order":
                shape = _data_dims_swapper(tuple(shape), value, ["detY", "angles", "detX"])
            elif key == "power_of_2_oversampling":
                power_of_2_oversampling = value
            elif key == "power_of_2_cropping":
                power_of_2_cropping = value
            elif key == "padding" and isinstance(value, int) and value >= 0:
                padding = value
        nz, nproj, data_n = (int(v) for v in shape)
        itemsize = int(np.dtype(kwargs.get("data_dtype", np.float32)).itemsize)
        recon_size = self.recon_size
        if recon_size > data_n:
            raise ValueError(
                "The reconstruction size {} should not be larger than the size of the horizontal detector {}".format(
                    recon_size, data_n
                )
            )
        stack.malloc(nz * nproj * data_n * itemsize)  # the caller's projections stay alive throughout
        odd_horiz, odd_vert = bool(data_n % 2), bool(nz % 2)
        raw_nz = nz
        data_n += odd_horiz
        nz += odd_vert
        padded = nz * nproj * data_n * 4 if (odd_horiz or odd_vert) else 0
        if padded:
            stack.malloc(padded)
        n = data_n + self.detectors_x_pad * 2 + padding * 2
        if power_of_2_cropping:
            n_pow2 = 2 ** math.ceil(math.log2(n))
            if 0.9 < n / n_pow2:
                n = n_pow2
        nz2 = nz // 2
        for b in (nproj * 4, nproj * 4, nproj * 8, nproj * 4):  # theta, sorted theta, int64 / int32 indices
            stack.malloc(b)
        stack.free(nproj * 8)
        tmp_p = self._fbp_filtering_estimator(data_n, n, nproj, nz, power_of_2_oversampling, oversampling_level)
        if padded:
            stack.free(padded)                        # `del data`: only our padded copy goes away
        datac, piece = self._setup_backprojection_input_estimator(n, nproj, nz2, tmp_p)
        recon_shape, recon = self.unpad_reconstructed_data_estimator(n, raw_nz, odd_horiz, recon_size)
        # STEPS 2-4 run per chunk of complex slices; every chunk has the same peak, so one is replayed
        self._fft_and_interpolation_estimator(piece)
        self.ifft_gathered_projections_estimator(piece)
        stack.free(piece)                             # the chunk's transform, after its unpadding
        stack.free(datac)
        stack.free(recon)                             # (like the reference, the result itself is not kept on the stack)
        for b in (nproj * 4, nproj * 4, nproj * 4):
            stack.free(b)
        return recon_shape

    # ---- the stage estimators (same names as the reference's; each returns the sizes the next stage needs) ----
    def _fbp_filtering_estimator(self, raw_width, width, nproj, nz, power_of_2_oversampling=True, oversampling_level=4):
        """STEP 0 (_fourier_filter; reference twin methodsDIR_CuPy.py:547-643): `out`, then per slice chunk the
        edge-padded rows, their rfft, the filtered spectrum and the irfft output plus one work area per transform.
        Returns the bytes of the filtered projections, which stay allocated."""
        from tomobar_b200.supp.memory_estimator_helpers import DeviceMemStack

        stack = DeviceMemStack.instance()
        if power_of_2_oversampling:
            over = 2 ** math.ceil(math.log2(raw_width * 3))
            if width > over:
                over = 2 ** math.ceil(math.log2(width))
        else:
            over = max(int(oversampling_level * raw_width), width)
        tmp_p = nz * nproj * width * 4  # (= the bytes of the complex slice pairs the chunks are packed into)
        stack.malloc(tmp_p)
        per = max(1, (1 << 27) // (nproj * over))
        if per > 1:
            per -= per % 2  # whole slice pairs per chunk
        rows_real, rows_cplx = min(per, nz) * nproj * over * 4, min(per, nz) * nproj * (over // 2 + 1) * 8
        if per % 2 == 0 and nz % 2 == 0 and over % 2 == 0 and self._FILTER_SLICE_PAIRS:
            # complex slice-pair rows: per / 2 * nproj * over * 8 bytes = rows_real
            stack.malloc(over * 8)                             # the two-sided filter
            stack.malloc(rows_real)                            # edge-padded slice pairs
            stack.malloc(rows_real), stack.malloc(rows_real)   # their fft + its work area
            stack.free(rows_real), stack.free(rows_real)       # (work area, then the padded rows)
            stack.malloc(rows_real), stack.malloc(rows_real)   # inverse transform + its work area
            stack.free(rows_real), stack.free(rows_real), stack.free(rows_real), stack.free(over * 8)
            return tmp_p
        stack.malloc(rows_real)                            # edge-padded rows
        stack.malloc(rows_cplx), stack.malloc(rows_cplx)   # rfft output + its work area
        stack.free(rows_cplx)
        stack.malloc(rows_cplx)                            # filtered spectrum
        stack.free(rows_cplx)                              # (the rfft output is released after the product)
        stack.malloc(rows_real), stack.malloc(rows_real)   # irfft output + its work area
        stack.free(rows_real), stack.free(rows_cplx), stack.free(rows_real), stack.free(rows_real)
        return tmp_p

    def _setup_backprojection_input_estimator(self, n, nproj, nz2, tmp_p):
        """STEP 1 (reference twin :685-699): the complex slice pairs replace the filtered projections; their 1-D
        FFT runs out of place (output + work area).  Returns the bytes of datac and of one chunk of the oversampled
        grid (STEPS 2-4 run chunk by chunk of complex slices)."""
        from tomobar_b200.supp.memory_estimator_helpers import DeviceMemStack

        stack = DeviceMemStack.instance()
        datac = nz2 * nproj * n * 8                    # == tmp_p: the filter stage packed its chunks into it
        assert datac == tmp_p
        stack.malloc(datac), stack.malloc(datac)      # FFT output + work area
        stack.free(datac), stack.free(datac)
        chunk = max(1, min(nz2, (1 << 28) // (4 * n * n)))
        if self._GATHER_SLICE_PAIRS and nz2 % 8 == 0 and chunk % 8 == 0:
            stack.malloc(datac)                       # the scale / sign pass writes the slice-pair layout out of place
            stack.free(datac)                         # (whole-grid branch, the default centre size)
        return datac, chunk * (2 * n) * (2 * n) * 8

    def _fft_and_interpolation_estimator(self, piece):
        """STEP 2 (reference twin :837-849): one chunk of the oversampled Cartesian grid is allocated and gathered /
        scattered into.  No angle-range table here, whatever the centre size: the gather finds its ranges on the fly."""
        from tomobar_b200.supp.memory_estimator_helpers import DeviceMemStack

        DeviceMemStack.instance().malloc(piece)

    def ifft_gathered_projections_estimator(self, piece):
        """STEP 3 (reference twin :898-918): the chunk's inverse 2-D FFT runs out of place (output + work area), then
        its input is released."""
        from tomobar_b200.supp.memory_estimator_helpers import DeviceMemStack

        stack = DeviceMemStack.instance()
        stack.malloc(piece), stack.malloc(piece)
        stack.free(piece), stack.free(piece)

    def unpad_reconstructed_data_estimator(self, n, raw_nz, odd_horiz, recon_size):
        """STEP 4 (reference twin :968-989): the reconstruction, allocated before the chunk loop.  Returns its shape and
        bytes (the caller releases it: like the reference, the result itself is not kept on the stack)."""
        from tomobar_b200.supp.memory_estimator_helpers import DeviceMemStack

        odd_recon = bool(recon_size % 2)
        um = (n - odd_horiz) // 2 - recon_size // 2
        up = (n - odd_horiz) // 2 + (recon_size + odd_recon) // 2
        recon_shape = (raw_nz, up - um, up - um)
        recon = int(np.prod(recon_shape)) * 4
        DeviceMemStack.instance().malloc(recon)
        return recon_shape, recon

    def _fourier_filter(self, data, raw_width, width, power_of_2_oversampling, oversampling_level, filter_type,
                        cutoff_freq, pack_into=None):
        """rfft -> analytic filter with the rotation-axis phase ramp -> irfft on an edge-padded,
        oversampled detector; cropped to ``width`` (methodsDIR_CuPy.py:449-545).

        ``pack_into`` (the complex slice pairs ``datac[nz/2][nproj][width]`` of FOURIER_INV's next step): every chunk
        of slices is cropped AND packed (r2c_c1dfftshift, fft_us_kernels.cu:529-557) straight out of the oversampled
        ``irfft`` output -- the filtered projections are never written out as an array of their own.  Returns None then."""
        if power_of_2_oversampling:
            over = 2 ** math.ceil(math.log2(raw_width * 3))
            if width > over:
                over = 2 ** math.ceil(math.log2(width))
        else:
            over = max(int(oversampling_level * raw_width), width)
        padding_m = over // 2 - raw_width // 2
        unpad_m = over // 2 - width // 2
        unpad_p = over // 2 + width // 2
        rotation_axis = self.centre_of_rotation + 0.5
        dev = data.device
        wfilter = torch.as_tensor(calc_filter(over, filter_type, cutoff_freq), device=dev)
        t = torch.fft.rfftfreq(over, device=dev).to(torch.float32)
        w = wfilter * torch.exp((-2 * np.pi * 1j * rotation_axis) * t.to(torch.complex64))
        if over % 2 == 0:
            # irfft is defined to ignore the imaginary part of the Nyquist bin, which the phase ramp makes non-zero
            # (purely imaginary for a centred axis); cuFFT's c2r result is undefined for such input and measurably
            # uses it at 8192-point rows (7.7e-3 relative against a float64 irfft at config 4) -- the bin is made real
            torch.view_as_real(w)[over // 2, 1] = 0.0
        nz, nproj, _ = data.shape
        # slice chunks bound the oversampled temporaries (the reference chunks for the same reason)
        per = max(1, (1 << 27) // (nproj * over))
        if pack_into is not None and per > 1:
            per -= per % 2  # whole slice pairs per chunk
        fused = pack_into is not None and per % 2 == 0 and nz % 2 == 0
        st = torch.cuda.current_stream(dev).cuda_stream
        if fused and over % 2 == 0 and self._FILTER_SLICE_PAIRS:
            # Slice PAIRS are filtered as complex rows: the filter's impulse response is real (the Nyquist bin of the
            # two-sided spectrum is the real one above), hence real and imaginary part are filtered independently -- one c2c transform pair per slice pair instead of an r2c / c2r pair per
            # slice, no c2r input clone, and the 1 / over of the inverse transform folded into the filter.
            h = over // 2
            wfull = torch.empty(over, dtype=torch.complex64, device=dev)
            wfull[:h] = w[:h]
            wfull[h] = w[h]
            wfull[h + 1:] = torch.conj(w[1:h]).flip(0)
            wfull *= 1.0 / over
            data = data.contiguous()
            for z0 in range(0, nz, per):
                cnt = min(per, nz - z0) // 2
                tmp = torch.empty((cnt, nproj, over), dtype=torch.complex64, device=dev)
                check(lib.tmb_edge_pad_pair(ptr(data[z0:]), ptr(tmp), cnt, nproj, raw_width, over, padding_m, st),
                      "tmb_edge_pad_pair")
                tmp = torch.fft.fft(tmp, dim=2)
                tmp.mul_(wfull)
                tmp = torch.fft.ifft(tmp, dim=2, norm="forward")
                check(lib.tmb_fi_crop_sign(ptr(tmp) + 8 * unpad_m, over, ptr(pack_into[z0 // 2:]), width, cnt * nproj, st),
                      "tmb_fi_crop_sign")
            return None
        out = None if fused else torch.empty((nz, nproj, width), dtype=torch.float32, device=dev)
        for z0 in range(0, nz, per):
            tmp = edge_pad(data[z0:z0 + per], padding_m, raw_width + 2 * padding_m)
            tmp = torch.fft.irfft(w * torch.fft.rfft(tmp, dim=2), n=over, dim=2)
            if fused:  # crop and pack straight out of the oversampled irfft output
                check(lib.tmb_fi_pack_rows(ptr(tmp) + 4 * unpad_m, over, nproj * over, ptr(pack_into[z0 // 2:]), width, nproj,
                                           tmp.shape[0] // 2, st), "tmb_fi_pack_rows")
            else:
                out[z0:z0 + per] = tmp[:, :, unpad_m:unpad_p]
        if fused:
            return None
        if pack_into is not None:
            check(lib.tmb_fi_pack(ptr(out), ptr(pack_into), width, nproj, nz // 2, st), "tmb_fi_pack")
            return None
        return out
