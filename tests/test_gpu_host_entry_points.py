"""tmb_fp3d_host / tmb_bp3d_host (include/tmb.h: host pointers in, host pointers out -- the calls a caller without
device arrays binds) against the oracle and against the device-pointer entry points; and a single
process that reconstructs on cuda:0 and then on cuda:1 (the shared-memory attribute of the kernels is per device)."""

import ctypes as C

import numpy as np
import pytest
import torch

from conftest import rel_max

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nz,n,nu,na,os_n", [(5, 48, 64, 30, None), (33, 70, 70, 40, 4)])
def test_host_entry_points_vs_oracle(oracle, nz, n, nu, na, os_n):
    from tomobar_b200._lib import lib, check
    from tomobar_b200.projector import ProjTools3D

    rng = np.random.default_rng(nz)
    angles = np.linspace(0, np.pi, na, endpoint=False).astype(np.float32)
    P = ProjTools3D(nu, 0, nz, angles, 0.25, n, "gpu", 0, os_n)
    O = oracle.Atools(nu, 0, nz, angles, 0.25, n, os_n)
    fp = C.c_void_p  # plain host addresses
    vol = rng.standard_normal((nz, n, n)).astype(np.float32)
    for sub in ([-1] if os_n is None else [0, os_n - 1]):
        na_s = na if sub < 0 else len(O.tbl_os[sub])
        sino = np.empty((nz, na_s, nu), np.float32)
        check(lib.tmb_fp3d_host(P._g, sub, vol.ctypes.data_as(fp), sino.ctypes.data_as(fp)), "tmb_fp3d_host")
        want = O._forwprojCuPy(vol) if sub < 0 else O._forwprojOSCuPy(vol, sub)
        assert rel_max(sino, want) < 2e-6
        dev = P._forwprojCuPy(torch.from_numpy(vol).cuda()) if sub < 0 else P._forwprojOSCuPy(torch.from_numpy(vol).cuda(), sub)
        assert np.array_equal(sino, dev.cpu().numpy())  # same kernels behind both entry points
        back = np.empty((nz, n, n), np.float32)
        check(lib.tmb_bp3d_host(P._g, sub, want.ctypes.data_as(fp), back.ctypes.data_as(fp)), "tmb_bp3d_host")
        want_b = O._backprojCuPy(want) if sub < 0 else O._backprojOSCuPy(want, sub)
        assert rel_max(back, want_b) < 2e-6


def test_one_process_two_devices(oracle):
    """cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device; it used to be set once per process,
    so the second device of a process failed every FP / TV launch."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from tomobar_b200.methodsIR_CuPy import RecToolsIRCuPy

    nz, n, na = 20, 64, 48
    angles = np.linspace(0, np.pi, na, endpoint=False).astype(np.float32)
    rng = np.random.default_rng(0)
    b = rng.random((nz, na, n)).astype(np.float32)
    res = []
    for d in (0, 1):
        rec = RecToolsIRCuPy(n, 0, nz, 0.0, angles, n, d, 4)
        x = rec.FISTA({"projection_data": torch.from_numpy(b).to(f"cuda:{d}")},
                      {"iterations": 2, "lipschitz_const": 3000.0, "nonnegativity": True},
                      {"method": "PD_TV", "regul_param": 1e-3, "iterations": 6})
        assert x.device.index == d
        res.append(x.cpu().numpy())
    assert np.array_equal(res[0], res[1])


def test_two_geometries_on_two_streams(oracle):
    """the per-angle constants of ALL geometries share one __constant__ table per device.  Two geometries
    used alternately on two streams must not overwrite the table under each other's kernels (uploads wait, on the
    device, for the streams still reading the resident table)."""
    from tomobar_b200.projector import ProjTools3D

    nz, n = 24, 96
    geoms = [ProjTools3D(n, 0, nz, np.linspace(0, np.pi, na, endpoint=False).astype(np.float32), cor, n, "gpu", 0, None)
             for na, cor in ((90, 0.0), (75, 1.5))]
    g = torch.Generator(device="cuda").manual_seed(5)
    vols = [torch.rand((nz, n, n), device="cuda", generator=g) for _ in geoms]
    want = [G._backprojCuPy(G._forwprojCuPy(v)).clone() for G, v in zip(geoms, vols)]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    got = [[], []]
    for _ in range(6):
        for i, (G, v, s) in enumerate(zip(geoms, vols, streams)):
            with torch.cuda.stream(s):
                got[i].append(G._backprojCuPy(G._forwprojCuPy(v)))
    torch.cuda.synchronize()
    for i in range(2):
        for r in got[i]:
            assert torch.equal(r, want[i])


@pytest.mark.parametrize("shape,pad_left,width_out", [((3, 7, 64), 32, 128), ((2, 5, 100), 6, 112), ((4, 3, 37), 5, 51),
                                                      ((1, 9, 48), 0, 48), ((2, 4, 33), 7, 44), ((1, 1, 1), 3, 8)])
def test_edge_pad_kernels(shape, pad_left, width_out):
    """supp.suppTools.edge_pad (suppTools.py:425-459 of the reference: np.pad(..., mode='edge') of the detector axis)
    through both kernels: 128-bit stores when the output rows are whole float4s, the scalar kernel otherwise."""
    from tomobar_b200.supp.suppTools import edge_pad

    g = torch.Generator(device="cuda").manual_seed(width_out)
    x = torch.randn(shape, device="cuda", generator=g)
    y = edge_pad(x, pad_left, width_out)
    ref = np.pad(x.cpu().numpy(), ((0, 0), (0, 0), (pad_left, width_out - pad_left - shape[-1])), mode="edge")
    assert y.shape == ref.shape
    np.testing.assert_array_equal(y.cpu().numpy(), ref)


@pytest.mark.parametrize("shape,pad_left,width_out", [((4, 7, 64), 32, 128), ((2, 5, 100), 6, 112), ((6, 3, 37), 5, 52),
                                                      ((2, 9, 48), 0, 48), ((2, 1, 1), 3, 8)])
def test_edge_pad_pair_and_crop_sign(shape, pad_left, width_out):
    """The two kernels either side of FOURIER_INV's complex slice-pair filter: tmb_edge_pad_pair (np.pad(mode='edge') of
    slices 2t and 2t+1 into the real / imaginary part of one complex row) and tmb_fi_crop_sign (crop of
    methodsDIR_CuPy.py:541-545 plus the (-1)^(x+1) of r2c_c1dfftshift, fft_us_kernels.cu:529-557).  Bit-exact."""
    from tomobar_b200._lib import check, lib
    from tomobar_b200._tensors import ptr

    g = torch.Generator(device="cuda").manual_seed(width_out)
    x = torch.randn(shape, device="cuda", generator=g)
    nz, rows, w = shape
    st = torch.cuda.current_stream().cuda_stream
    y = torch.empty((nz // 2, rows, width_out), dtype=torch.complex64, device="cuda")
    check(lib.tmb_edge_pad_pair(ptr(x), ptr(y), nz // 2, rows, w, width_out, pad_left, st), "tmb_edge_pad_pair")
    ref = np.pad(x.cpu().numpy(), ((0, 0), (0, 0), (pad_left, width_out - pad_left - w)), mode="edge")
    got = y.cpu().numpy()
    np.testing.assert_array_equal(got.real, ref[0::2])
    np.testing.assert_array_equal(got.imag, ref[1::2])
    n = max(1, width_out // 2)
    off = (width_out - n) // 2
    z = torch.empty((nz // 2, rows, n), dtype=torch.complex64, device="cuda")
    check(lib.tmb_fi_crop_sign(ptr(y) + 8 * off, width_out, ptr(z), n, (nz // 2) * rows, st), "tmb_fi_crop_sign")
    sgn = np.where(np.arange(n) % 2 == 1, 1.0, -1.0).astype(np.float32)
    np.testing.assert_array_equal(z.cpu().numpy(), got[:, :, off:off + n] * sgn)
