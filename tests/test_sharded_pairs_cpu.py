"""End-to-end CPU run of ``ShardedPDTV(pairs=True)``: the real host code (slab layout, peer-pointer
arithmetic, pairing of iterations, odd tail, ping-pong) drives the real CUDA source of the z-shard kernel
(k_pd_tv3d_f2s<GHOST> under tests/warp_shim) on three "ranks" that are threads sharing fake symmetric
memory (numpy buffers).  The assembled volume must equal N plain whole-volume iterations.

What is faked: the symmetric-memory slab (numpy), the neighbour synchronisation (a thread barrier), the
two C entry points (tmb_pd_tv_iter2 -> the warp shim, tmb_pd_tv_iter -> numpy).  This path has not run on
GPUs yet; this test is its stand-in until round 2."""

import contextlib
import ctypes as C
import importlib.util
import os
import threading
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from test_warp_shim_fused_tv import FP, F32, _aligned, shim  # noqa: F401  (fixture re-export)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def plain():
    spec = importlib.util.spec_from_file_location("emulate_pd_fused2", os.path.join(ROOT, "tools", "emulate_pd_fused2.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _scalars(lam, lip):
    tau = F32(np.float64(lam) * 0.1)
    sigma = F32(1.0 / (np.float64(lip) * np.float64(tau)))
    lt = F32(np.float64(tau) / np.float64(lam))
    return sigma, tau, lt, F32(1.0)


def _view(addr, shape):
    n = int(np.prod(shape))
    return np.ctypeslib.as_array(C.cast(int(addr), FP), shape=(n,)).reshape(shape)


@pytest.mark.parametrize("nz_total,world,iters,methodTV,nonneg", [(14, 3, 7, 0, 1), (12, 2, 4, 1, 0), (10, 3, 5, 0, 0)])
def test_sharded_pairs_end_to_end_on_the_cpu(shim, plain, monkeypatch, nz_total, world, iters, methodTV, nonneg):
    import tomobar_b200._lib as tlib
    import tomobar_b200._tensors as ttens
    import tomobar_b200.zshard as zs

    ny, nx, lam, lip = 9, 132, 4e-2, 12.0
    plane = ny * nx
    sigma, tau, lt, theta = _scalars(lam, lip)
    shim_lock, barrier = threading.Lock(), threading.Barrier(world)
    slabs = {}

    class FakeSlab:
        def __init__(self, nbytes, device, group):
            self.rank = device  # the test passes the rank where the device goes
            self.np = _aligned((int(nbytes) // 4,), 0.0)
            self.buf = torch.from_numpy(self.np.view(np.uint8))
            slabs[self.rank] = self
            barrier.wait()  # every rank has allocated: publish the "peer mappings"
            self.ptrs = [slabs[r].np.ctypes.data for r in range(world)]

        def view(self, offset, shape, dtype):
            assert dtype == torch.float32 and offset % 4 == 0
            n = int(np.prod(shape))
            return torch.from_numpy(self.np[offset // 4:offset // 4 + n].reshape(shape))

        def barrier(self):
            barrier.wait()

    class FakeSync:
        def __init__(self, slab, shard, mode):
            pass

        def acquire(self):
            barrier.wait()

        def produced(self):
            pass

    def fake_iter2(inp, u_in, u_out, p1i, p2i, p3i, p1o, p2o, p3o, dz, dy, dx, lam_, method, nn, lip_, glo, ghi, *rest):
        ghosts, s = rest[:10], _scalars(lam_, lip_)
        cast = lambda a: C.cast(int(a), FP) if a else None  # noqa: E731
        with shim_lock:  # the shim keeps its shared memory in one global array
            variant = 11 if not p1i else 3  # no dual inputs: the first pair of a prox call (PZERO)
            rc = shim.shim_run_fused_tv(variant, int(nn), int(method), *[cast(a) for a in (inp, u_in, u_out, p1i, p2i, p3i, p1o, p2o, p3o)],
                                        *s, dx, dy, dz, max(1, (dz + 1) // 2), int(glo), int(ghi), *[cast(g) for g in ghosts])
        return rc

    def fake_iter(inp, u_in, u_out, p1i, p2i, p3i, p1o, p2o, p3o, dz, dy, dx, lam_, method, nn, lip_, half, glo, ghi,
                  u_lo, p1_lo, p2_lo, p3_lo, u_hi, stream):
        """One plain iteration of a shard: the ghost planes are stacked around it, the plain whole-volume
        iteration runs on the stack and the shard's own planes are kept."""
        s = _scalars(lam_, lip_)
        shp = (dz, dy, dx)
        U, P = _view(u_in, shp), [_view(p, shp) for p in (p1i, p2i, p3i)]
        D = _view(inp, shp)
        lo, hi = (1 if glo else 0), (1 if ghi else 0)
        one = (1, dy, dx)
        Ue = np.concatenate(([_view(u_lo, one)] if lo else []) + [U] + ([_view(u_hi, one)] if hi else []))
        Pe = [np.concatenate(([_view(pl, one)] if lo else []) + [p] + ([np.zeros(one, F32)] if hi else []))
              for p, pl in zip(P, (p1_lo, p2_lo, p3_lo))]
        De = np.concatenate(([np.zeros(one, F32)] if lo else []) + [D] + ([np.zeros(one, F32)] if hi else []))
        Un, Pn = plain.iterate_plain(De, Ue, Pe, *s, bool(nn), bool(method))
        _view(u_out, shp)[...] = Un[lo:lo + dz]
        for dst, src in zip((p1o, p2o, p3o), Pn):
            _view(dst, shp)[...] = src[lo:lo + dz]
        return 0

    fake_lib = SimpleNamespace(tmb_pd_tv_iter2=fake_iter2, tmb_pd_tv_iter=fake_iter, tmb_last_error=lambda: b"")
    monkeypatch.setattr(tlib, "lib", fake_lib)
    monkeypatch.setattr(ttens, "stream_ptr", lambda t: 0)
    monkeypatch.setattr(zs, "_PeerSlab", FakeSlab)
    monkeypatch.setattr(zs, "_NeighbourSync", FakeSync)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())

    rng = np.random.default_rng(nz_total * 10 + world)
    vol = (0.2 * rng.standard_normal((nz_total, ny, nx))).astype(F32)
    results, errors = {}, []

    def rank_main(rank):
        try:
            z0, z1 = zs.shard_bounds(nz_total, world, rank, 2)
            shard = SimpleNamespace(nz_total=nz_total, world=world, multiple=2, rank=rank, group=None, z0=z0, z1=z1,
                                    nz_local=z1 - z0, prev=rank - 1 if rank > 0 else None,
                                    next=rank + 1 if rank + 1 < world else None, _global=lambda p: p,
                                        require_tv_shards=lambda what="": None)
            tv = zs.ShardedPDTV(shard, (z1 - z0, ny, nx), rank, False, peer_memory=True, sync="signals", pairs=True)
            assert tv.pairs
            data = torch.from_numpy(np.ascontiguousarray(vol[z0:z1]))
            for _ in range(2):  # the buffers are reused across calls
                results[rank] = tv(data, lam, iters, methodTV, nonneg, lip).numpy().copy()
        except BaseException as e:  # noqa: BLE001
            errors.append((rank, repr(e)))
            barrier.abort()

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors

    U, P = vol.copy(), [np.zeros_like(vol) for _ in range(3)]
    for _ in range(iters):
        U, P = plain.iterate_plain(vol, U, P, sigma, tau, lt, theta, bool(nonneg), bool(methodTV))
    got = np.concatenate([results[r] for r in range(world)], axis=0)
    assert np.isfinite(got).all()
    assert np.max(np.abs(got - U)) <= 2e-6 * max(np.max(np.abs(U)), 1.0)
