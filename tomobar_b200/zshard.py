"""z-sharding of the reconstruction over ``torch.distributed`` ranks (one process per GPU).

For a scalar centre of rotation and a vertical rotation axis every slice is an independent 2-D
problem for the projector pair, the FBP filter and FOURIER_INV (the reference relies on the caller,
HTTomo, to chunk in z: docs/source/introduction/dependencies.rst:56-58).  Rank r owns the contiguous
block of slices ``[z0, z1)`` of the volume and the same detector rows of the sinogram.  What couples
the shards:

* scalars: the power method's norm (methodsIR_CuPy.py:333-351), the PWLS weight normalisation
  ``w.max()`` (:394-395), CGLS inner products (:270-289)  ->  one scalar all-reduce each;
* 3-D total variation: the forward z difference and the backward z divergence reach one plane into
  the neighbouring shards (primal_dual_for_total_variation.cu:188-194, 244-252).  ``ShardedPDTV`` /
  ``ShardedROFTV`` let the TV kernel read the neighbours' boundary planes directly over NVLink
  (buffers in symmetric memory, peer pointers, one cross-GPU barrier per inner iteration) or, as a
  fallback, refresh ghost planes with point-to-point messages; either way the result is
  bit-identical to the whole-volume prox;
* the final volume: one all-gather (``ZShard.all_gather_volume``).

Nothing here touches the CUDA library except the two ``Sharded*TV`` classes; the rest runs on any backend
(``gloo`` on CPU tensors in the unit tests, ``nccl`` on the GPUs).
"""

from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(nz: int, world: int, rank: int, multiple: int = 2) -> Tuple[int, int]:
    """Contiguous block ``[z0, z1)`` of rank ``rank``.  Block sizes are multiples of ``multiple``
    (FOURIER_INV processes slices in pairs, methodsDIR_CuPy.py:268-282) except the last non-empty one."""
    if nz <= 0 or world <= 0 or not 0 <= rank < world or multiple <= 0:
        raise ValueError("shard_bounds: bad arguments")
    per = -(-nz // world)
    per = -(-per // multiple) * multiple
    z0 = min(nz, rank * per)
    return z0, min(nz, z0 + per)


def shard_table(nz: int, world: int, multiple: int = 2) -> List[Tuple[int, int]]:
    """The blocks of all ranks.  Raises when a rank would own no slices -- the same error on EVERY rank, because
    every rank evaluates the same table (one rank raising alone would leave the others waiting in a collective)."""
    bounds = [shard_bounds(nz, world, r, multiple) for r in range(world)]
    empty = [r for r, (a, b) in enumerate(bounds) if b <= a]
    if empty:
        raise ValueError(f"ranks {empty} of {world} would own no slices of {nz}: use fewer ranks")
    return bounds


class ZShard:
    """The z-partition of one rank and the collectives the hot path needs."""

    def __init__(self, nz_total: int, group: Optional[dist.ProcessGroup] = None, multiple: int = 2):
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        self.nz_total = int(nz_total)
        self.multiple = multiple
        # every rank knows every rank's block: decisions that must be the same everywhere (is a rank empty,
        # can the 3-D TV prox run across the shards) are taken from this table, never from the local shape
        self.bounds = shard_table(self.nz_total, self.world, multiple)
        self.sizes = [b[1] - b[0] for b in self.bounds]
        self.z0, self.z1 = self.bounds[self.rank]
        # neighbours that own slices (trailing ranks may be empty only if the constructor raised there)
        self.prev = self.rank - 1 if self.rank > 0 else None
        self.next = self.rank + 1 if self.rank + 1 < self.world and self.z1 < self.nz_total else None

    @property
    def nz_local(self) -> int:
        return self.z1 - self.z0

    @property
    def min_size(self) -> int:
        """Slices of the smallest shard (the same number on every rank)."""
        return min(self.sizes)

    def require_tv_shards(self, what: str = "3-D TV") -> None:
        """The sharded TV kernels need >= 2 planes in EVERY shard (an odd ``nz_total`` can leave the last rank
        with one).  The check uses the global table, so all ranks raise together instead of one rank taking a
        different path and the others waiting for it in a rendezvous / semaphore."""
        if self.world > 1 and self.min_size < 2:
            raise ValueError(f"{what} across z-shards needs at least two slices per rank; {self.nz_total} slices "
                             f"over {self.world} ranks give blocks of {self.sizes}: use fewer ranks or an even "
                             "number of slices")

    def _global(self, peer: int) -> int:
        return dist.get_global_rank(self.group, peer) if self.group is not None else peer

    # ---- scalars ------------------------------------------------------------------------------
    def allreduce_(self, t: torch.Tensor, op=None) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM if op is None else op, group=self.group)
        return t

    def norm(self, x: torch.Tensor) -> torch.Tensor:
        """Euclidean norm of the WHOLE (sharded) array; 0-dim tensor on x's device."""
        s = torch.sum(x.double() * x.double())
        return torch.sqrt(self.allreduce_(s)).to(torch.float32)

    def dot(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        return self.allreduce_(torch.sum(a.double() * b.double())).to(torch.float32)

    def max(self, x: torch.Tensor) -> torch.Tensor:
        return self.allreduce_(x.max().clone(), dist.ReduceOp.MAX)

    def min(self, x: torch.Tensor) -> torch.Tensor:
        return self.allreduce_(x.min().clone(), dist.ReduceOp.MIN)

    # ---- halos ----------------------------------------------------------------------------------
    def exchange_halos(self, up: Sequence[Tuple[torch.Tensor, torch.Tensor]],
                       down: Sequence[Tuple[torch.Tensor, torch.Tensor]]) -> None:
        """``up``: pairs (send, recv): ``send`` goes to the next rank, ``recv`` is filled by the
        previous rank's ``send`` of the same pair.  ``down``: ``send`` goes to the previous rank,
        ``recv`` is filled by the next rank.  All tensors contiguous; blocks until complete (on the
        current CUDA stream for nccl)."""
        if self.world == 1:
            return
        ops: List[dist.P2POp] = []
        for send, recv in up:
            if self.next is not None:
                ops.append(dist.P2POp(dist.isend, send, self._global(self.next), self.group))
            if self.prev is not None:
                ops.append(dist.P2POp(dist.irecv, recv, self._global(self.prev), self.group))
        for send, recv in down:
            if self.prev is not None:
                ops.append(dist.P2POp(dist.isend, send, self._global(self.prev), self.group))
            if self.next is not None:
                ops.append(dist.P2POp(dist.irecv, recv, self._global(self.next), self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    # ---- assembly -------------------------------------------------------------------------------
    def all_gather_volume(self, x_local: torch.Tensor) -> torch.Tensor:
        """All ranks receive the whole volume ``[nz_total, ...]`` (the single collective of the
        projector / FBP / FOURIER_INV paths)."""
        if self.world == 1:
            return x_local
        per = shard_bounds(self.nz_total, self.world, 0, self.multiple)[1]
        tail = tuple(x_local.shape[1:])
        padded = x_local
        if x_local.shape[0] != per:  # last shard: pad to the common block size
            padded = torch.zeros((per,) + tail, dtype=x_local.dtype, device=x_local.device)
            padded[: x_local.shape[0]] = x_local
        out = torch.empty((self.world * per,) + tail, dtype=x_local.dtype, device=x_local.device)
        dist.all_gather_into_tensor(out, padded.contiguous(), group=self.group)
        return out[: self.nz_total]


class _PeerSlab:
    """One buffer per rank, allocated from torch's symmetric memory and mapped into every peer's
    address space over NVLink / NVSwitch (``torch.distributed._symmetric_memory``).  The layout is
    the same on all ranks, so a neighbour's array lives at ``buffer_ptrs[neighbour] + offset``."""

    def __init__(self, nbytes: int, device: torch.device, group: Optional[dist.ProcessGroup]):
        import torch.distributed._symmetric_memory as symm_mem

        self.buf = symm_mem.empty(int(nbytes), dtype=torch.uint8, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.ptrs = list(self.hdl.buffer_ptrs)

    def view(self, offset: int, shape, dtype) -> torch.Tensor:
        n = int(torch.empty((), dtype=dtype).element_size())
        for d in shape:
            n *= int(d)
        return self.buf[offset:offset + n].view(dtype).view(*shape)

    def barrier(self) -> None:
        """Cross-GPU barrier on the current stream (signal pads in peer memory)."""
        self.hdl.barrier(0)

    # pairwise stream-ordered semaphores (one per ordered pair of ranks and channel)
    def signal(self, peers) -> None:
        for r in peers:
            self.hdl.put_signal(int(r), 1)

    def wait(self, peers) -> None:
        for r in peers:
            self.hdl.wait_signal(int(r), 1)


class _NeighbourSync:
    """Ordering of the peer-memory TV iterations.  Before a rank launches iteration k it needs its two
    neighbours to have finished iteration k - 1 (they wrote the planes it is about to read and read the
    planes it is about to overwrite).  ``mode="signals"``: pairwise semaphores with the two neighbours
    (a put after every kernel, a wait before the next one); ``mode="barrier"``: one barrier over all
    ranks per iteration."""

    def __init__(self, slab: _PeerSlab, shard: "ZShard", mode: str):
        if mode not in ("signals", "barrier"):
            raise ValueError("sync mode must be 'signals' or 'barrier'")
        self.slab, self.mode = slab, mode
        self.peers = [shard._global(r) for r in (shard.prev, shard.next) if r is not None]
        self.pending = False  # neighbours have signalled the end of their previous work

    def produced(self) -> None:
        """This rank's buffers are ready for the neighbours / it no longer reads theirs."""
        if self.mode == "signals":
            self.slab.signal(self.peers)
            self.pending = True

    def acquire(self) -> None:
        """Wait until the neighbours have done the same."""
        if self.mode == "barrier":
            self.slab.barrier()
        elif self.pending:
            self.slab.wait(self.peers)
            self.pending = False


def _peer_memory_default(shard: "ZShard", device: torch.device) -> bool:
    return shard.world > 1 and device.type == "cuda" and dist.get_backend(shard.group) == "nccl"


class ShardedPDTV:
    """PD_TV prox of a z-sharded 3-D volume, bit-identical to ``PD_TV_cupy`` on the whole volume.

    Two ways of getting the one-plane halos (top plane of U and P1..P3 of the previous shard, bottom
    plane of U of the next one):

    * ``peer_memory=True`` (default on NCCL): the ping-pong buffers live in symmetric memory and the
      kernel reads the neighbours' planes **directly over NVLink** through peer pointers
      (``tmb_pd_tv_iter(..., u_lo, p*_lo, u_hi)``): compute and halo transfer are one kernel, the
      host only places a cross-GPU barrier between iterations;
    * ``peer_memory=False``: ghost planes next to the shard, refreshed with point-to-point messages
      (5 planes per rank per iteration) before each launch.

    ``pairs`` (default with peer memory and fp32 duals; ``pairs=False`` or ``TMB_SHARDED_PAIRS=0`` turn it
    off): two iterations per pass through ``tmb_pd_tv_iter2`` -- the kernel reaches two planes into each
    neighbour (U, P1..P3 and the prox input, which then lives in symmetric memory too) and the neighbours
    synchronise once per PAIR of iterations.  Same arithmetic as the whole-volume prox, which pairs its
    iterations in the same kernel (tests/test_gpu_tv_shards.py, tests/test_gpu_multi.py).

    Buffers are allocated once and reused across calls."""

    def __init__(self, shard: ZShard, shape: Tuple[int, int, int], device: torch.device, half_precision: bool = False,
                 peer_memory: Optional[bool] = None, sync: str = "signals", pairs: Optional[bool] = None):
        nzl, ny, nx = shape
        if nzl != shard.nz_local:
            raise ValueError("ShardedPDTV: the volume shard does not match the z-partition")
        self.shard, self.shape, self.device, self.half = shard, (nzl, ny, nx), device, bool(half_precision)
        self.peer = _peer_memory_default(shard, device) if peer_memory is None else bool(peer_memory)
        shard.require_tv_shards("PD_TV")
        if pairs is None:
            pairs = os.environ.get("TMB_SHARDED_PAIRS", "1") != "0"
        # pairs of iterations: peer memory, fp32 duals, rows of whole float4s (every shard has >= 2 planes)
        self.pairs = bool(pairs) and self.peer and not self.half and nx % 4 == 0 and ny >= 2
        pdt = torch.float16 if self.half else torch.float32
        # U: ghost plane below (index 0) and above (index nzl + 1); P: ghost plane below only
        if not self.peer:
            self.U = [torch.zeros((nzl + 2, ny, nx), dtype=torch.float32, device=device) for _ in range(2)]
            self.P = [[torch.zeros((nzl + 1, ny, nx), dtype=pdt, device=device) for _ in range(3)] for _ in range(2)]
            return
        per = shard_bounds(shard.nz_total, shard.world, 0, shard.multiple)[1]  # largest shard: common layout
        plane = ny * nx
        esz = 2 if self.half else 4
        self._ub, self._pb = (per + 2) * plane * 4, (per + 1) * plane * esz
        self._plane, self._esz = plane, esz
        self._db = 2 * self._ub + 6 * self._pb  # offset of the prox input (pairs only)
        self.slab = _PeerSlab(self._db + (per * plane * 4 if self.pairs else 0), device, shard.group)
        self.slab.buf.zero_()
        self.slab.barrier()
        self.sync = _NeighbourSync(self.slab, shard, sync)
        self.U = [self.slab.view(a * self._ub, (nzl + 2, ny, nx), torch.float32) for a in range(2)]
        self.P = [[self.slab.view(2 * self._ub + (a * 3 + c) * self._pb, (nzl + 1, ny, nx), pdt) for c in range(3)]
                  for a in range(2)]
        self.D = self.slab.view(self._db, (nzl, ny, nx), torch.float32) if self.pairs else None

    def _ghost_ptrs(self, a: int):
        """Peer addresses of the halo planes of ping-pong set ``a``."""
        sh, plane, esz = self.shard, self._plane, self._esz
        u_lo = p_lo = u_hi = None
        if sh.prev is not None:
            base = self.slab.ptrs[sh._global(sh.prev)]
            z0p, z1p = shard_bounds(sh.nz_total, sh.world, sh.prev, sh.multiple)
            top = z1p - z0p  # index of the previous shard's last own plane (its plane 0 is a ghost)
            u_lo = base + a * self._ub + top * plane * 4
            p_lo = [base + 2 * self._ub + (a * 3 + c) * self._pb + top * plane * esz for c in range(3)]
        if sh.next is not None:
            u_hi = self.slab.ptrs[sh._global(sh.next)] + a * self._ub + plane * 4  # its first own plane
        return u_lo, p_lo, u_hi

    def _ghost_ptrs2(self, a: int, first: bool = False):
        """Peer addresses a fused pass needs from ping-pong set ``a``: the neighbours' last / first TWO
        planes of U, two / one planes of P1..P3 and one plane of the prox input (10 pointers in the
        argument order of ``tmb_pd_tv_iter2``).  ``first``: the first pair of a prox call reads the prox input
        as its primal variable (the neighbours' inputs as its ghost planes) and no dual variable at all."""
        sh, plane = self.shard, self._plane
        lo = [None] * 5
        hi = [None] * 5
        if sh.prev is not None:
            base = self.slab.ptrs[sh._global(sh.prev)]
            z0p, z1p = shard_bounds(sh.nz_total, sh.world, sh.prev, sh.multiple)
            top = z1p - z0p  # its own planes sit at indices 1 .. top of U and P, 0 .. top - 1 of the input
            if first:
                lo = [base + self._db + (top - 2) * plane * 4, None, None, None]
            else:
                lo = [base + a * self._ub + (top - 1) * plane * 4]
                lo += [base + 2 * self._ub + (a * 3 + c) * self._pb + (top - 1) * plane * 4 for c in range(3)]
            lo += [base + self._db + (top - 1) * plane * 4]
        if sh.next is not None:
            base = self.slab.ptrs[sh._global(sh.next)]
            if first:
                hi = [base + self._db, None, None, None]
            else:
                hi = [base + a * self._ub + plane * 4]
                hi += [base + 2 * self._ub + (a * 3 + c) * self._pb + plane * 4 for c in range(3)]
            hi += [base + self._db]
        return lo + hi

    def __call__(self, data: torch.Tensor, regularisation_parameter: float, iterations: int, methodTV: int = 0,
                 nonneg: int = 0, lipschitz_const: float = 8.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        from tomobar_b200._lib import lib, check
        from tomobar_b200._tensors import ptr, stream_ptr

        sh = self.shard
        nzl, ny, nx = self.shape
        if tuple(data.shape) != self.shape or data.dtype != torch.float32 or not data.is_contiguous():
            raise ValueError(f"ShardedPDTV: expected a contiguous float32 volume shard of shape {self.shape}")
        U, P = self.U, self.P
        if self.peer:
            self.sync.acquire()  # nobody still reads the buffers of the previous call
        ghost_lo, ghost_hi = int(sh.prev is not None), int(sh.next is not None)
        if self.pairs:
            self.D.copy_(data)  # the neighbours read planes of the prox input as well
            if int(iterations) < 2:  # a lone iteration goes through the one-iteration kernel on set 0
                U[0][1:nzl + 1].copy_(data)
                for c in range(3):
                    P[0][c].zero_()
            self.sync.produced()
            return self._run_pairs(data, regularisation_parameter, int(iterations), methodTV, nonneg, lipschitz_const,
                                   ghost_lo, ghost_hi, out)
        U[0][1:nzl + 1].copy_(data)
        for c in range(3):
            P[0][c].zero_()
        if self.peer:
            self.sync.produced()
        with torch.cuda.device(self.device):
            for it in range(int(iterations)):
                a, b = it % 2, 1 - it % 2
                u_lo = p_lo = u_hi = None
                if self.peer:
                    # the neighbours have finished writing set `a` (and reading set `b`): a pairwise
                    # semaphore (or one barrier) per iteration replaces the halo messages, the kernel
                    # loads the planes over NVLink
                    self.sync.acquire()
                    u_lo, p_lo, u_hi = self._ghost_ptrs(a)
                else:
                    up = [(U[a][nzl], U[a][0])]
                    if it > 0:  # the dual variable starts at zero everywhere
                        up += [(P[a][c][nzl], P[a][c][0]) for c in range(3)]
                    sh.exchange_halos(up, [(U[a][1], U[a][nzl + 1])])
                p_lo = p_lo or [None, None, None]
                check(lib.tmb_pd_tv_iter(ptr(data), ptr(U[a][1:]), ptr(U[b][1:]), ptr(P[a][0][1:]), ptr(P[a][1][1:]),
                                         ptr(P[a][2][1:]), ptr(P[b][0][1:]), ptr(P[b][1][1:]), ptr(P[b][2][1:]),
                                         nzl, ny, nx, float(regularisation_parameter), int(methodTV), int(nonneg),
                                         float(lipschitz_const), int(self.half), ghost_lo, ghost_hi,
                                         u_lo, p_lo[0], p_lo[1], p_lo[2], u_hi, stream_ptr(data)), "tmb_pd_tv_iter")
                if self.peer:
                    self.sync.produced()
        res = U[int(iterations) % 2][1:nzl + 1]
        if out is None:
            return res.clone()
        out.copy_(res)
        return out


    def _run_pairs(self, data, lam, iterations, methodTV, nonneg, lip, ghost_lo, ghost_hi, out):
        """Pairs of iterations per pass (peer memory): one neighbour synchronisation per launch.  The first pair
        reads the prox input (in symmetric memory, ``self.D``) as its primal variable and no dual variable --
        no copy into the primal buffer, no memsets -- and the last launch writes its primal result straight into
        ``out`` (nobody reads it over NVLink)."""
        from tomobar_b200._lib import lib, check
        from tomobar_b200._tensors import ptr, stream_ptr

        nzl, ny, nx = self.shape
        U, P, D = self.U, self.P, self.D
        if out is None:
            out = torch.empty_like(data)
        if iterations <= 0:
            out.copy_(data)
            return out
        it, a = 0, 0
        with torch.cuda.device(self.device):
            while it < iterations:
                b = 1 - a
                self.sync.acquire()  # the neighbours have finished writing set `a` and reading set `b`
                pair = it + 2 <= iterations
                last = it + (2 if pair else 1) >= iterations
                u_out = ptr(out) if last else ptr(U[b][1:])
                if pair:
                    first = it == 0
                    g = self._ghost_ptrs2(a, first)
                    u_in = ptr(D) if first else ptr(U[a][1:])
                    p_in = [None, None, None] if first else [ptr(P[a][c][1:]) for c in range(3)]
                    check(lib.tmb_pd_tv_iter2(ptr(D), u_in, u_out, p_in[0], p_in[1], p_in[2],
                                              ptr(P[b][0][1:]), ptr(P[b][1][1:]), ptr(P[b][2][1:]),
                                              nzl, ny, nx, float(lam), int(methodTV), int(nonneg), float(lip),
                                              ghost_lo, ghost_hi, *g, stream_ptr(data)), "tmb_pd_tv_iter2")
                    it += 2
                else:
                    u_lo, p_lo, u_hi = self._ghost_ptrs(a)
                    p_lo = p_lo or [None, None, None]
                    check(lib.tmb_pd_tv_iter(ptr(D), ptr(U[a][1:]), u_out, ptr(P[a][0][1:]), ptr(P[a][1][1:]),
                                             ptr(P[a][2][1:]), ptr(P[b][0][1:]), ptr(P[b][1][1:]), ptr(P[b][2][1:]),
                                             nzl, ny, nx, float(lam), int(methodTV), int(nonneg), float(lip), 0,
                                             ghost_lo, ghost_hi, u_lo, p_lo[0], p_lo[1], p_lo[2], u_hi,
                                             stream_ptr(data)), "tmb_pd_tv_iter")
                    it += 1
                self.sync.produced()
                a = b
        return out


class ShardedROFTV:
    """ROF_TV prox of a z-sharded 3-D volume, bit-identical to ``ROF_TV_cupy`` on the whole volume.

    The normalised z difference of the plane below a shard enters the divergence at its first
    plane (rudin_osher_fatemi_total_variation.cu:170-181, 235), and that difference itself needs
    the plane below it: two ghost planes below, one above -- read over NVLink from the neighbours'
    buffers (``peer_memory=True``, one cross-GPU barrier per iteration) or refreshed with messages
    (top two planes to the next rank, bottom plane to the previous one)."""

    def __init__(self, shard: ZShard, shape: Tuple[int, int, int], device: torch.device, half_precision: bool = False,
                 peer_memory: Optional[bool] = None, sync: str = "signals"):
        nzl, ny, nx = shape
        if nzl != shard.nz_local:
            raise ValueError("ShardedROFTV: the volume shard does not match the z-partition")
        shard.require_tv_shards("ROF_TV")
        self.shard, self.shape, self.device, self.half = shard, (nzl, ny, nx), device, bool(half_precision)
        self.peer = _peer_memory_default(shard, device) if peer_memory is None else bool(peer_memory)
        if not self.peer:
            self.U = [torch.zeros((nzl + 3, ny, nx), dtype=torch.float32, device=device) for _ in range(2)]
            return
        per = shard_bounds(shard.nz_total, shard.world, 0, shard.multiple)[1]
        self._plane = ny * nx
        self._ub = (per + 3) * self._plane * 4
        self.slab = _PeerSlab(2 * self._ub, device, shard.group)
        self.slab.buf.zero_()
        self.slab.barrier()
        self.sync = _NeighbourSync(self.slab, shard, sync)
        self.U = [self.slab.view(a * self._ub, (nzl + 3, ny, nx), torch.float32) for a in range(2)]

    def _ghost_ptrs(self, a: int):
        sh, plane = self.shard, self._plane
        u_lo = u_hi = None
        if sh.prev is not None:
            z0p, z1p = shard_bounds(sh.nz_total, sh.world, sh.prev, sh.multiple)
            # the previous shard's last two own planes (its planes 0 and 1 are ghosts)
            u_lo = self.slab.ptrs[sh._global(sh.prev)] + a * self._ub + (z1p - z0p) * plane * 4
        if sh.next is not None:
            u_hi = self.slab.ptrs[sh._global(sh.next)] + a * self._ub + 2 * plane * 4
        return u_lo, u_hi

    def __call__(self, data: torch.Tensor, regularisation_parameter: float, iterations: int,
                 time_marching_parameter: float = 0.001, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        from tomobar_b200._lib import lib, check
        from tomobar_b200._tensors import ptr, stream_ptr

        sh = self.shard
        nzl, ny, nx = self.shape
        if tuple(data.shape) != self.shape or data.dtype != torch.float32 or not data.is_contiguous():
            raise ValueError(f"ShardedROFTV: expected a contiguous float32 volume shard of shape {self.shape}")
        U = self.U
        if self.peer:
            self.sync.acquire()
        U[0][2:nzl + 2].copy_(data)
        if self.peer:
            self.sync.produced()
        ghost_lo, ghost_hi = int(sh.prev is not None), int(sh.next is not None)
        with torch.cuda.device(self.device):
            for it in range(int(iterations)):
                a, b = it % 2, 1 - it % 2
                u_lo = u_hi = None
                if self.peer:
                    self.sync.acquire()
                    u_lo, u_hi = self._ghost_ptrs(a)
                else:
                    sh.exchange_halos([(U[a][nzl:nzl + 2], U[a][0:2])], [(U[a][2], U[a][nzl + 2])])
                check(lib.tmb_rof_tv_iter(ptr(data), ptr(U[a][2:]), ptr(U[b][2:]), nzl, ny, nx,
                                          float(regularisation_parameter), float(time_marching_parameter),
                                          int(self.half), ghost_lo, ghost_hi, u_lo, u_hi, stream_ptr(data)),
                          "tmb_rof_tv_iter")
                if self.peer:
                    self.sync.produced()
        res = U[int(iterations) % 2][2:nzl + 2]
        if out is None:
            return res.clone()
        out.copy_(res)
        return out
