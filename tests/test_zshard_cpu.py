"""Host logic of the z-sharded path on CPU: partition arithmetic, and a world_size-2 ``gloo`` run
of the collectives the hot path uses (whole-volume norm / dot / max, halo exchange, final
all-gather, the sharded power method on the oracle's operators)."""

import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tomobar_b200.zshard import ZShard, shard_bounds


@pytest.mark.parametrize("nz,world", [(512, 8), (512, 1), (10, 2), (7, 4), (384, 4), (130, 8), (2, 2)])
def test_shard_bounds_cover_and_even(nz, world):
    blocks = [shard_bounds(nz, world, r) for r in range(world)]
    assert blocks[0][0] == 0 and max(b[1] for b in blocks) == nz
    for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
        assert a1 == b0 and a0 <= a1  # contiguous, ordered
    sizes = [b[1] - b[0] for b in blocks if b[1] > b[0]]
    assert all(s % 2 == 0 for s in sizes[:-1])  # slice pairs stay together (FOURIER_INV)
    assert sum(sizes) == nz


def test_shard_bounds_errors():
    with pytest.raises(ValueError):
        shard_bounds(0, 2, 0)
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def test_single_process_shard_is_identity():
    sh = ZShard(12)
    assert (sh.z0, sh.z1, sh.world, sh.prev, sh.next) == (0, 12, 1, None, None)
    x = torch.arange(24.0).reshape(12, 2)
    assert torch.equal(sh.all_gather_volume(x), x)
    assert float(sh.norm(x)) == pytest.approx(float(torch.linalg.vector_norm(x)))
    sh.exchange_halos([(x[0], x[1])], [])  # no-op


def _worker(rank, world, init_file, nz, results):
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    try:
        sh = ZShard(nz)
        rng = np.random.default_rng(5)
        full = torch.from_numpy(rng.standard_normal((nz, 6, 8)).astype(np.float32))
        mine = full[sh.z0:sh.z1].clone()
        out = {"bounds": (sh.z0, sh.z1)}
        out["norm"] = float(sh.norm(mine))
        out["dot"] = float(sh.dot(mine, 2 * mine))
        out["max"] = float(sh.max(mine))
        out["min"] = float(sh.min(mine))
        # halo exchange: ghost planes below / above
        buf = torch.full((sh.nz_local + 2, 6, 8), -7.0)
        buf[1:-1] = mine
        sh.exchange_halos([(buf[sh.nz_local], buf[0])], [(buf[1], buf[sh.nz_local + 1])])
        lo_ok = torch.equal(buf[0], full[sh.z0 - 1]) if sh.prev is not None else bool((buf[0] == -7).all())
        hi_ok = torch.equal(buf[-1], full[sh.z1]) if sh.next is not None else bool((buf[-1] == -7).all())
        out["halo_ok"] = bool(lo_ok and hi_ok)
        out["gather_ok"] = bool(torch.equal(sh.all_gather_volume(mine), full))

        # sharded power method on the oracle's operators == whole-volume power method
        from oracle import oracle as O

        O.build()
        n, na = 24, 20
        angles = np.linspace(0, np.pi, na, endpoint=False).astype(np.float32)
        rec_loc = O.RecIR(n, 0, sh.nz_local, 0.0, angles, n, None)
        x = rng.standard_normal((nz, n, n)).astype(np.float32)[sh.z0:sh.z1]
        s = None
        for _ in range(6):
            y = rec_loc._Ax(x)
            x = rec_loc._Atb(y)
            s = float(sh.norm(torch.from_numpy(x)))
            x = x / np.float32(s)
        out["L"] = s
        results[rank] = out
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_collectives_and_power_method():
    nz, world = 10, 2
    with tempfile.TemporaryDirectory() as d:
        mgr = mp.Manager()
        results = mgr.dict()
        mp.spawn(_worker, args=(world, os.path.join(d, "rdzv"), nz, results), nprocs=world, join=True)
    rng = np.random.default_rng(5)
    full = rng.standard_normal((nz, 6, 8)).astype(np.float32)
    assert results[0]["bounds"] == (0, 6) and results[1]["bounds"] == (6, 10)
    for r in range(world):
        assert results[r]["norm"] == pytest.approx(float(np.linalg.norm(full.astype(np.float64))), rel=1e-6)
        assert results[r]["dot"] == pytest.approx(2 * float(np.sum(full.astype(np.float64) ** 2)), rel=1e-6)
        assert results[r]["max"] == float(full.max()) and results[r]["min"] == float(full.min())
        assert results[r]["halo_ok"] and results[r]["gather_ok"]
    # whole-volume power method in one process (same starting vector)
    from oracle import oracle as O

    O.build()
    n, na = 24, 20
    angles = np.linspace(0, np.pi, na, endpoint=False).astype(np.float32)
    rec = O.RecIR(n, 0, nz, 0.0, angles, n, None)
    x = rng.standard_normal((nz, n, n)).astype(np.float32)
    s = None
    for _ in range(6):
        x = rec._Atb(rec._Ax(x))
        s = float(np.linalg.norm(x.astype(np.float64)))
        x = x / np.float32(s)
    assert results[0]["L"] == pytest.approx(s, rel=1e-5)
    assert results[0]["L"] == results[1]["L"]


def test_peer_pointers_of_a_fused_pass():
    """ShardedPDTV._ghost_ptrs2: where a two-iterations-per-pass launch looks for its ghost planes inside the
    neighbours' symmetric-memory slabs (host arithmetic only: the slab layout is U[2] | P[2][3] | input)."""
    from types import SimpleNamespace

    from tomobar_b200.zshard import ShardedPDTV

    nz_total, world, ny, nx = 22, 3, 6, 8          # shards of 8, 8 and 6 planes
    plane = ny * nx
    per = shard_bounds(nz_total, world, 0, 2)[1]
    ub, pb = (per + 2) * plane * 4, (per + 1) * plane * 4
    bases = [1 << 40, 2 << 40, 3 << 40]

    def make(rank):
        tv = object.__new__(ShardedPDTV)
        tv.shard = SimpleNamespace(nz_total=nz_total, world=world, multiple=2, rank=rank,
                                   prev=rank - 1 if rank > 0 else None, next=rank + 1 if rank + 1 < world else None,
                                   _global=lambda peer: peer)
        tv._plane, tv._esz, tv._ub, tv._pb, tv._db = plane, 4, ub, pb, 2 * ub + 6 * pb
        tv.slab = SimpleNamespace(ptrs=bases)
        return tv

    for a in (0, 1):
        g = make(1)._ghost_ptrs2(a)                # the middle shard has both neighbours
        lo_n = 8                                   # planes of shard 0: its own planes are U[1..8], P[1..8], input[0..7]
        assert g[0] == bases[0] + a * ub + (lo_n - 1) * plane * 4                              # U planes -2, -1
        assert g[1:4] == [bases[0] + 2 * ub + (a * 3 + c) * pb + (lo_n - 1) * plane * 4 for c in range(3)]
        assert g[4] == bases[0] + 2 * ub + 6 * pb + (lo_n - 1) * plane * 4                      # input plane -1
        assert g[5] == bases[2] + a * ub + plane * 4                                           # U planes dz, dz+1
        assert g[6:9] == [bases[2] + 2 * ub + (a * 3 + c) * pb + plane * 4 for c in range(3)]
        assert g[9] == bases[2] + 2 * ub + 6 * pb                                              # input plane dz
        first, last = make(0)._ghost_ptrs2(a), make(2)._ghost_ptrs2(a)
        assert first[:5] == [None] * 5 and all(p is not None for p in first[5:])
        assert last[5:] == [None] * 5 and all(p is not None for p in last[:5])
        # U_lo must address the same bytes as the single-iteration ghost pointer, one plane earlier
        u_lo1, p_lo1, u_hi1 = make(1)._ghost_ptrs(a)
        assert g[0] == u_lo1 - plane * 4 and g[5] == u_hi1
        assert g[1:4] == [p - plane * 4 for p in p_lo1]


def test_shard_decisions_are_collective():
    """What must be the same on every rank comes from the table of ALL blocks (an odd slice count left
    the last rank with one slice and on a different code path than its neighbours, which then waited forever)."""
    from tomobar_b200.zshard import ZShard, shard_table

    with pytest.raises(ValueError, match="own no slices"):
        shard_table(6, 4)  # blocks of 2, 2, 2, 0: every rank raises, not only rank 3
    assert shard_table(5, 2) == [(0, 4), (4, 5)]
    # the one-slice last block: every rank refuses the sharded 3-D TV with the same error
    for rank in range(2):
        sh = ZShard.__new__(ZShard)
        sh.rank, sh.world, sh.nz_total, sh.multiple = rank, 2, 5, 2
        sh.bounds = shard_table(5, 2)
        sh.sizes = [b - a for a, b in sh.bounds]
        assert sh.min_size == 1
        with pytest.raises(ValueError, match="at least two slices per rank"):
            sh.require_tv_shards("PD_TV")
    sh.world, sh.nz_total, sh.bounds = 2, 8, shard_table(8, 2)
    sh.sizes = [4, 4]
    sh.require_tv_shards()
