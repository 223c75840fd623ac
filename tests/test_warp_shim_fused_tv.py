"""The CUDA source of the fused PD_TV kernels (tomobar_b200/csrc/tmb_tv_fused.cuh), compiled UNCHANGED with
g++ under tests/warp_shim (one OS thread per lane, shuffles as barrier exchanges, NaN-poisoned shared
memory) and run on the CPU against two plain whole-volume iterations.

This is how the kernel variants that have not run on a GPU yet are checked at source level: the z-shard
(GHOST) instantiation behind tmb_pd_tv_iter2 with its peer-pointer arithmetic, and the four-CTAs-per-SM
variant; the two variants that HAVE run on the B200 go through the same harness as its control.
Tolerance 2e-6 (the host has no MUFU unit and contracts differently); logic errors show up as O(1e-2)."""

import ctypes as C
import importlib.util
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "tests", "warp_shim")
F32 = np.float32
FP = C.POINTER(C.c_float)


@pytest.fixture(scope="module")
def shim():
    out_dir = os.path.join(SHIM, "_build")
    os.makedirs(out_dir, exist_ok=True)
    lib_path = os.path.join(out_dir, "libshim_fused_tv.so")
    srcs = [os.path.join(SHIM, "run_fused_tv.cpp"), os.path.join(SHIM, "cuda_shim.h"),
            os.path.join(ROOT, "tomobar_b200", "csrc", "tmb_tv_fused.cuh")]
    if not os.path.exists(lib_path) or any(os.path.getmtime(s) > os.path.getmtime(lib_path) for s in srcs):
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
                               srcs[0], "-o", lib_path], cwd=SHIM)
    lib = C.CDLL(lib_path)
    lib.shim_run_fused_tv.restype = C.c_int
    lib.shim_run_fused_tv.argtypes = ([C.c_int] * 3 + [FP] * 9 + [C.c_float] * 4 + [C.c_int] * 6 + [FP] * 10)
    return lib


@pytest.fixture(scope="module")
def plain():
    spec = importlib.util.spec_from_file_location("emulate_pd_fused2", os.path.join(ROOT, "tools", "emulate_pd_fused2.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _ptr(a):
    return a.ctypes.data_as(FP) if a is not None else None


def _aligned(shape, fill=None):
    """float32 array on a 64-byte boundary (the kernels use 128-bit accesses)."""
    n = int(np.prod(shape))
    raw = np.empty(n * 4 + 64, dtype=np.uint8)
    off = (-raw.ctypes.data) % 64
    a = raw[off:off + n * 4].view(F32).reshape(shape)
    if fill is not None:
        a[...] = fill
    return a


def _case(shape, seed):
    rng = np.random.default_rng(seed)
    inp = _aligned(shape, rng.standard_normal(shape).astype(F32))
    U = _aligned(shape, (inp + 0.3 * rng.standard_normal(shape)).astype(F32))
    P = [_aligned(shape, (0.7 * rng.standard_normal(shape)).astype(F32)) for _ in range(3)]
    return inp, U, P


SIGMA, TAU, LT, THETA = F32(0.9), F32(0.05), F32(0.37), F32(1.0)


def _two_plain(plain, inp, U, P, nonneg, aniso):
    U1, P1 = plain.iterate_plain(inp, U, P, SIGMA, TAU, LT, THETA, nonneg, aniso)
    return plain.iterate_plain(inp, U1, P1, SIGMA, TAU, LT, THETA, nonneg, aniso)


def _close(a, b):
    assert np.isfinite(a).all()
    assert np.max(np.abs(a - b)) <= 2e-6 * max(np.max(np.abs(b)), 1.0)


@pytest.mark.parametrize("variant", [0, 1, 2, 4, 6, 7, 10, 12, 14])
@pytest.mark.parametrize("shape,zrun,nonneg,aniso", [
    ((2, 3, 8), 2, False, False),
    ((5, 9, 124), 5, True, False),
    ((7, 18, 132), 3, False, False),
    ((6, 5, 244), 2, True, True),
    ((4, 5, 4), 1, False, False),
    ((5, 45, 12), 2, True, False),   # several strips of 8 rows along y, the last one partial
])
def test_whole_volume_kernels_on_the_cpu(shim, plain, variant, shape, zrun, nonneg, aniso):
    """0: k_pd_tv3d_f2 (the default, validated on the B200), 1: k_pd_tv3d_f2s (run on the B200 on five
    shapes), 2: k_pd_tv3d_f2s at four CTAs per SM, 4: with packets two rows ahead, 6: with L2 prefetches,
    12 / 14: strips of 8 rows (14: packets two rows ahead),
    7 / 10: k_pd_tv3d_f2t, the TMA-fed variant with a ring of 4 / 8 stages (bulk copies as memcpy, mbarrier waits as
    warp barriers; lanes outside the volume read NaN-poisoned stage memory and must not reach a stored lane)."""
    inp, U, P = _case(shape, sum(shape))
    U2, P2 = _two_plain(plain, inp, U, P, nonneg, aniso)
    Uo = _aligned(shape, np.nan)
    Q = [_aligned(shape, np.nan) for _ in range(3)]
    dz, dy, dx = shape
    rc = shim.shim_run_fused_tv(variant, int(nonneg), int(aniso), _ptr(inp), _ptr(U), _ptr(Uo), *[_ptr(p) for p in P],
                                *[_ptr(q) for q in Q], SIGMA, TAU, LT, THETA, dx, dy, dz, zrun, 0, 0, *([None] * 10))
    assert rc == 0
    _close(Uo, U2)
    for c in range(3):
        _close(Q[c], P2[c])


@pytest.mark.parametrize("shape,cuts,zrun,nonneg,aniso", [
    ((8, 9, 124), [4], 4, True, False),
    ((9, 6, 12), [2, 5], 2, False, False),      # shards of 2, 3 and 4 planes
    ((10, 18, 132), [3, 7], 8, False, True),
    ((6, 5, 8), [2, 4], 1, True, False),        # every z-run starts in the neighbour's planes
])
@pytest.mark.parametrize("variant", [3, 8])
def test_z_shard_kernel_with_peer_pointers_on_the_cpu(shim, plain, variant, shape, cuts, zrun, nonneg, aniso):
    """k_pd_tv3d_f2s<GHOST> exactly as tmb_pd_tv_iter2 launches it: every shard is its own set of arrays and
    the ghost pointers aim into the NEIGHBOURS' arrays (their last two / first two planes), like the peer
    mappings of ShardedPDTV(pairs=True).  Assembled result == two plain iterations of the whole volume."""
    inp, U, P = _case(shape, sum(shape) + 1)
    U2, P2 = _two_plain(plain, inp, U, P, nonneg, aniso)
    bounds = list(zip([0] + list(cuts), list(cuts) + [shape[0]]))
    _, dy, dx = shape
    pl = dy * dx

    def own(a, z0, z1):  # a shard's own copy (separate allocation, like another GPU's memory)
        return _aligned((z1 - z0, dy, dx), a[z0:z1])

    S = [dict(inp=own(inp, a, b), U=own(U, a, b), P=[own(p, a, b) for p in P], n=b - a) for a, b in bounds]
    outs = []
    for i, s in enumerate(S):
        lo = S[i - 1] if i > 0 else None
        hi = S[i + 1] if i + 1 < len(S) else None
        ghost = [None] * 10
        if lo is not None:  # planes -2, -1 of U and P: the neighbour's last two; plane -1 of the input
            ghost[0] = _ptr(lo["U"][lo["n"] - 2:])
            ghost[1:4] = [_ptr(p[lo["n"] - 2:]) for p in lo["P"]]
            ghost[4] = _ptr(lo["inp"][lo["n"] - 1:])
        if hi is not None:  # planes dz, dz + 1 of U, plane dz of P and the input: the neighbour's first
            ghost[5] = _ptr(hi["U"])
            ghost[6:9] = [_ptr(p) for p in hi["P"]]
            ghost[9] = _ptr(hi["inp"])
        Uo = _aligned(s["U"].shape, np.nan)
        Q = [_aligned(s["U"].shape, np.nan) for _ in range(3)]
        rc = shim.shim_run_fused_tv(variant, int(nonneg), int(aniso), _ptr(s["inp"]), _ptr(s["U"]), _ptr(Uo),
                                    *[_ptr(p) for p in s["P"]], *[_ptr(q) for q in Q], SIGMA, TAU, LT, THETA, dx, dy,
                                    s["n"], zrun, int(lo is not None), int(hi is not None), *ghost)
        assert rc == 0
        outs.append((Uo, Q))
    _close(np.concatenate([o[0] for o in outs], axis=0), U2)
    for c in range(3):
        _close(np.concatenate([o[1][c] for o in outs], axis=0), P2[c])


@pytest.mark.parametrize("shape,zrun,nonneg,aniso", [((5, 9, 124), 5, True, False), ((7, 18, 132), 3, False, True),
                                                     ((4, 5, 4), 1, False, False), ((3, 40, 8), 3, True, False)])
@pytest.mark.parametrize("variant", [5, 9, 13])
def test_first_pass_variant_on_the_cpu(shim, plain, variant, shape, zrun, nonneg, aniso):
    """k_pd_tv3d_f2s<PZERO> (hook 6's first pass of a prox call): the dual arrays are NOT read -- they hold
    NaN here -- and the input doubles as the primal variable; result == two plain iterations from P = 0."""
    inp, _, _ = _case(shape, sum(shape) + 7)
    zeros = [np.zeros(shape, F32) for _ in range(3)]
    U2, P2 = _two_plain(plain, inp, inp, zeros, nonneg, aniso)
    poison = [_aligned(shape, np.nan) for _ in range(3)]
    Uo = _aligned(shape, np.nan)
    Q = [_aligned(shape, np.nan) for _ in range(3)]
    dz, dy, dx = shape
    rc = shim.shim_run_fused_tv(variant, int(nonneg), int(aniso), _ptr(inp), _ptr(inp), _ptr(Uo), *[_ptr(p) for p in poison],
                                *[_ptr(q) for q in Q], SIGMA, TAU, LT, THETA, dx, dy, dz, zrun, 0, 0, *([None] * 10))
    assert rc == 0
    _close(Uo, U2)
    for c in range(3):
        _close(Q[c], P2[c])

