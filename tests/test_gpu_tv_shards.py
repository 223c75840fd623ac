"""z-sharded 3-D PD_TV on ONE GPU: the volume is cut into z-blocks that are advanced in lock step
through ``tmb_pd_tv_iter`` with their ghost planes refreshed by plain copies between the inner
iterations (what ``ShardedPDTV`` does with point-to-point messages).  The assembled result must be
bit-identical to the whole-volume prox.  (The multi-process version: tests/test_gpu_multi.py.)"""

import os

import numpy as np
import pytest
import torch

from conftest import rel_max, single_iteration_tv

pytestmark = pytest.mark.gpu


def _vol(shape, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    v = torch.randn(shape, device="cuda", generator=g) * 0.01
    v += (torch.rand(shape, device="cuda", generator=g) > 0.5).float() * 0.02
    return v


def _sharded_prox(v, cuts, lam, iters, methodTV, nonneg, lip, half):
    from tomobar_b200._lib import lib, check
    from tomobar_b200._tensors import ptr, stream_ptr

    nz, ny, nx = v.shape
    bounds = list(zip([0] + cuts, cuts + [nz]))
    pdt = torch.float16 if half else torch.float32
    S = []
    for (z0, z1) in bounds:
        nzl = z1 - z0
        U = [torch.zeros((nzl + 2, ny, nx), device="cuda") for _ in range(2)]
        P = [[torch.zeros((nzl + 1, ny, nx), dtype=pdt, device="cuda") for _ in range(3)] for _ in range(2)]
        U[0][1:nzl + 1] = v[z0:z1]
        S.append(dict(nzl=nzl, U=U, P=P, data=v[z0:z1].contiguous()))
    for it in range(iters):
        a, b = it % 2, 1 - it % 2
        for i, s in enumerate(S):  # halo refresh (ShardedPDTV.exchange)
            if i + 1 < len(S):
                nxt = S[i + 1]
                nxt["U"][a][0].copy_(s["U"][a][s["nzl"]])
                for c in range(3):
                    nxt["P"][a][c][0].copy_(s["P"][a][c][s["nzl"]])
                s["U"][a][s["nzl"] + 1].copy_(nxt["U"][a][1])
        for i, s in enumerate(S):
            U, P, nzl = s["U"], s["P"], s["nzl"]
            check(lib.tmb_pd_tv_iter(ptr(s["data"]), ptr(U[a][1:]), ptr(U[b][1:]), ptr(P[a][0][1:]), ptr(P[a][1][1:]),
                                     ptr(P[a][2][1:]), ptr(P[b][0][1:]), ptr(P[b][1][1:]), ptr(P[b][2][1:]),
                                     nzl, ny, nx, lam, methodTV, nonneg, lip, int(half), int(i > 0),
                                     int(i + 1 < len(S)), None, None, None, None, None, stream_ptr(v)),
                  "tmb_pd_tv_iter")
    return torch.cat([s["U"][iters % 2][1:s["nzl"] + 1] for s in S], dim=0)


@pytest.mark.parametrize("shape,cuts", [((40, 36, 64), [20]), ((45, 21, 132), [8, 30]), ((70, 16, 260), [2, 36]),
                                        ((96, 8, 128), [48])])
@pytest.mark.parametrize("methodTV,nonneg", [(0, 1), (1, 0)])
@pytest.mark.parametrize("half", [False, True])
def test_sharded_pd_tv_is_bit_identical(shape, cuts, methodTV, nonneg, half):
    from tomobar_b200.regularisersCuPy import PD_TV_cupy

    v = _vol(shape, 11)
    with single_iteration_tv():
        whole = PD_TV_cupy(v, 4e-4, 7, methodTV, nonneg, 12.0, 0, half)
    parts = _sharded_prox(v, list(cuts), 4e-4, 7, methodTV, nonneg, 12.0, half)
    torch.cuda.synchronize()
    assert torch.equal(whole, parts)
    # the default whole-volume prox (pairs of iterations fused) agrees to rounding
    fused = PD_TV_cupy(v, 4e-4, 7, methodTV, nonneg, 12.0, 0, half)
    assert rel_max(fused.cpu().numpy(), whole.cpu().numpy()) < 2e-6
    # and independent blocks (the reference under HTTomo's z-chunking) do differ at the seams
    blocks = torch.cat([PD_TV_cupy(v[a:b].contiguous(), 4e-4, 7, methodTV, nonneg, 12.0, 0, half)
                        for a, b in zip([0] + list(cuts), list(cuts) + [shape[0]])], dim=0)
    assert not torch.equal(whole, blocks)


def _sharded_rof(v, cuts, lam, iters, tau, half):
    from tomobar_b200._lib import lib, check
    from tomobar_b200._tensors import ptr, stream_ptr

    nz, ny, nx = v.shape
    bounds = list(zip([0] + cuts, cuts + [nz]))
    S = []
    for (z0, z1) in bounds:
        nzl = z1 - z0
        U = [torch.zeros((nzl + 3, ny, nx), device="cuda") for _ in range(2)]
        U[0][2:nzl + 2] = v[z0:z1]
        S.append(dict(nzl=nzl, U=U, data=v[z0:z1].contiguous()))
    for it in range(iters):
        a, b = it % 2, 1 - it % 2
        for i, s in enumerate(S):
            if i + 1 < len(S):
                nxt = S[i + 1]
                nxt["U"][a][0:2].copy_(s["U"][a][s["nzl"]:s["nzl"] + 2])
                s["U"][a][s["nzl"] + 2].copy_(nxt["U"][a][2])
        for i, s in enumerate(S):
            U, nzl = s["U"], s["nzl"]
            check(lib.tmb_rof_tv_iter(ptr(s["data"]), ptr(U[a][2:]), ptr(U[b][2:]), nzl, ny, nx, lam, tau, int(half),
                                      int(i > 0), int(i + 1 < len(S)), None, None, stream_ptr(v)), "tmb_rof_tv_iter")
    return torch.cat([s["U"][iters % 2][2:s["nzl"] + 2] for s in S], dim=0)


@pytest.mark.parametrize("shape,cuts", [((40, 36, 64), [20]), ((45, 21, 132), [8, 30]), ((70, 16, 260), [2, 36]),
                                        ((96, 8, 128), [48])])
@pytest.mark.parametrize("half", [False, True])
def test_sharded_rof_tv_is_bit_identical(shape, cuts, half):
    from tomobar_b200.regularisersCuPy import ROF_TV_cupy

    v = _vol(shape, 12)
    whole = ROF_TV_cupy(v, 4e-4, 7, 1e-3, 0, half)
    parts = _sharded_rof(v, list(cuts), 4e-4, 7, 1e-3, half)
    torch.cuda.synchronize()
    assert torch.equal(whole, parts)


def test_ghost_planes_need_the_strip_kernel():
    from tomobar_b200._lib import lib
    from tomobar_b200._tensors import ptr

    v = _vol((6, 10, 30), 1)  # dx % 4 != 0
    U = torch.zeros((8, 10, 30), device="cuda")
    P = [torch.zeros((7, 10, 30), device="cuda") for _ in range(6)]
    rc = lib.tmb_pd_tv_iter(ptr(v), ptr(U[1:]), ptr(torch.zeros_like(U)[1:]), *[ptr(p[1:]) for p in P], 6, 10, 30,
                            1e-3, 0, 0, 12.0, 0, 1, 0, None, None, None, None, None, None)
    assert rc == -3 and b"ghost" in lib.tmb_last_error()


@pytest.mark.parametrize("half", [False, True])
def test_explicit_ghost_pointers(half):
    """The ghost planes may live anywhere (on the multi-GPU path they are the neighbour GPU's own
    buffers, read over NVLink): two shards of one volume, each reading the other's boundary planes
    in place through u_lo / p*_lo / u_hi -- no copies at all -- give the whole-volume result."""
    from tomobar_b200._lib import lib, check
    from tomobar_b200._tensors import ptr, stream_ptr
    from tomobar_b200.regularisersCuPy import PD_TV_cupy, ROF_TV_cupy

    v = _vol((40, 20, 64), 21)
    nz, ny, nx = v.shape
    cut, iters = 24, 6
    pdt = torch.float16 if half else torch.float32
    bounds = [(0, cut), (cut, nz)]
    # own planes only: no ghost slots
    U = [[torch.zeros((b - a, ny, nx), device="cuda") for (a, b) in bounds] for _ in range(2)]
    P = [[[torch.zeros((b - a, ny, nx), dtype=pdt, device="cuda") for _ in range(3)] for (a, b) in bounds]
         for _ in range(2)]
    D = [v[a:b].contiguous() for (a, b) in bounds]
    for i in range(2):
        U[0][i].copy_(D[i])
    for it in range(iters):
        a, b = it % 2, 1 - it % 2
        n0 = cut
        # shard 0: ghost above = first plane of shard 1; shard 1: ghosts below = last plane of shard 0
        check(lib.tmb_pd_tv_iter(ptr(D[0]), ptr(U[a][0]), ptr(U[b][0]), *[ptr(P[a][0][c]) for c in range(3)],
                                 *[ptr(P[b][0][c]) for c in range(3)], n0, ny, nx, 4e-4, 0, 1, 12.0, int(half), 0, 1,
                                 None, None, None, None, ptr(U[a][1][0]), stream_ptr(v)), "tmb_pd_tv_iter")
        check(lib.tmb_pd_tv_iter(ptr(D[1]), ptr(U[a][1]), ptr(U[b][1]), *[ptr(P[a][1][c]) for c in range(3)],
                                 *[ptr(P[b][1][c]) for c in range(3)], nz - cut, ny, nx, 4e-4, 0, 1, 12.0, int(half),
                                 1, 0, ptr(U[a][0][n0 - 1]), *[ptr(P[a][0][c][n0 - 1]) for c in range(3)], None,
                                 stream_ptr(v)), "tmb_pd_tv_iter")
    got = torch.cat([U[iters % 2][0], U[iters % 2][1]], dim=0)
    with single_iteration_tv():
        assert torch.equal(got, PD_TV_cupy(v, 4e-4, iters, 0, 1, 12.0, 0, half))

    R = [[torch.zeros((b - a, ny, nx), device="cuda") for (a, b) in bounds] for _ in range(2)]
    for i in range(2):
        R[0][i].copy_(D[i])
    for it in range(iters):
        a, b = it % 2, 1 - it % 2
        check(lib.tmb_rof_tv_iter(ptr(D[0]), ptr(R[a][0]), ptr(R[b][0]), cut, ny, nx, 4e-4, 1e-3, int(half), 0, 1,
                                  None, ptr(R[a][1][0]), stream_ptr(v)), "tmb_rof_tv_iter")
        check(lib.tmb_rof_tv_iter(ptr(D[1]), ptr(R[a][1]), ptr(R[b][1]), nz - cut, ny, nx, 4e-4, 1e-3, int(half), 1, 0,
                                  ptr(R[a][0][cut - 2]), None, stream_ptr(v)), "tmb_rof_tv_iter")
    got = torch.cat([R[iters % 2][0], R[iters % 2][1]], dim=0)
    assert torch.equal(got, ROF_TV_cupy(v, 4e-4, iters, 1e-3, 0, half))


# ---- pairs of iterations per pass (tmb_pd_tv_iter2) -------------------------------------------------
def _sharded_prox_pairs(v, cuts, lam, iters, methodTV, nonneg, lip, pzero_first=False):
    """z-blocks advanced in lock step, two iterations per pass: U keeps two ghost planes on either side,
    P two below and one above, Input one on either side (adjacent memory, refreshed by copies once per
    PAIR); an odd last iteration goes through tmb_pd_tv_iter on the same buffers."""
    from tomobar_b200._lib import lib, check
    from tomobar_b200._tensors import ptr, stream_ptr

    nz, ny, nx = v.shape
    bounds = list(zip([0] + cuts, cuts + [nz]))
    S = []
    for (z0, z1) in bounds:
        nzl = z1 - z0
        U = [torch.zeros((nzl + 4, ny, nx), device="cuda") for _ in range(2)]
        P = [[torch.zeros((nzl + 3, ny, nx), device="cuda") for _ in range(3)] for _ in range(2)]
        D = torch.zeros((nzl + 2, ny, nx), device="cuda")
        U[0][2:nzl + 2] = v[z0:z1]
        D[1:nzl + 1] = v[z0:z1]
        S.append(dict(nzl=nzl, U=U, P=P, D=D))
    for i, s in enumerate(S):  # Input halos never change
        if i + 1 < len(S):
            S[i + 1]["D"][0].copy_(s["D"][s["nzl"]])
            s["D"][s["nzl"] + 1].copy_(S[i + 1]["D"][1])

    def refresh(a):
        for i, s in enumerate(S):
            if i + 1 < len(S):
                nxt, n = S[i + 1], s["nzl"]
                nxt["U"][a][0:2].copy_(s["U"][a][n:n + 2])          # its planes -2, -1 = our last two
                s["U"][a][n + 2:n + 4].copy_(nxt["U"][a][2:4])       # our planes dz, dz+1 = its first two
                for c in range(3):
                    nxt["P"][a][c][0:2].copy_(s["P"][a][c][n:n + 2])
                    s["P"][a][c][n + 2].copy_(nxt["P"][a][c][2])

    it, a = 0, 0
    while it < iters:
        b = 1 - a
        refresh(a)
        pair = it + 2 <= iters
        for i, s in enumerate(S):
            U, P, D, nzl = s["U"], s["P"], s["D"], s["nzl"]
            lo, hi = int(i > 0), int(i + 1 < len(S))
            if pair:
                # the first pair of a prox call may be told that the dual variable is zero (it is then not read)
                p_in = [None] * 3 if (pzero_first and it == 0) else [ptr(P[a][c][2:]) for c in range(3)]
                check(lib.tmb_pd_tv_iter2(ptr(D[1:]), ptr(U[a][2:]), ptr(U[b][2:]), *p_in,
                                          *[ptr(P[b][c][2:]) for c in range(3)], nzl, ny, nx, lam, methodTV, nonneg,
                                          lip, lo, hi, *([None] * 10), stream_ptr(v)), "tmb_pd_tv_iter2")
            else:
                check(lib.tmb_pd_tv_iter(ptr(D[1:]), ptr(U[a][2:]), ptr(U[b][2:]), *[ptr(P[a][c][2:]) for c in range(3)],
                                         *[ptr(P[b][c][2:]) for c in range(3)], nzl, ny, nx, lam, methodTV, nonneg,
                                         lip, 0, lo, hi, None, None, None, None, None, stream_ptr(v)), "tmb_pd_tv_iter")
        it += 2 if pair else 1
        a = b
    return torch.cat([s["U"][a][2:s["nzl"] + 2] for s in S], dim=0)


@pytest.mark.parametrize("shape,cuts", [((40, 36, 64), [20]), ((45, 21, 132), [8, 30]), ((70, 16, 260), [2, 36]),
                                        ((96, 8, 128), [48])])
@pytest.mark.parametrize("methodTV,nonneg,iters", [(0, 1, 6), (1, 0, 7)])
def test_sharded_pairs_of_pd_iterations(shape, cuts, methodTV, nonneg, iters):
    from tomobar_b200.regularisersCuPy import PD_TV_cupy

    v = _vol(shape, 17)
    with single_iteration_tv():
        whole = PD_TV_cupy(v, 4e-4, iters, methodTV, nonneg, 12.0, 0, False)
    for pzero_first in (False, True):
        parts = _sharded_prox_pairs(v, list(cuts), 4e-4, iters, methodTV, nonneg, 12.0, pzero_first)
        torch.cuda.synchronize()
        assert rel_max(parts.cpu().numpy(), whole.cpu().numpy()) < 2e-6
