"""Two-GPU (one process per GPU, NCCL) checks of the z-sharded path.  Skipped on a single-GPU box;
run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""

import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")]


def _worker(rank, world, init_file, results):
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"file://{init_file}", rank=rank, world_size=world, device_id=dev)
    try:
        from tomobar_b200.methodsIR_CuPy import RecToolsIRCuPy
        from tomobar_b200.regularisersCuPy import PD_TV_cupy
        from tomobar_b200.regularisersCuPy import ROF_TV_cupy
        from conftest import single_iteration_tv  # whole-volume references through the kernel the shards launch
        from tomobar_b200.zshard import ShardedPDTV, ShardedROFTV, ZShard

        nz, n, na = 24, 64, 48
        sh = ZShard(nz)
        g = torch.Generator(device="cpu").manual_seed(3)
        full = (torch.randn((nz, n, n), generator=g) * 0.01 + (torch.rand((nz, n, n), generator=g) > 0.5) * 0.02)
        full = full.to(dev)
        out = {}
        # --- sharded PD_TV prox == whole-volume prox, bit for bit --------------------------------
        out["tv_equal_half0"] = out["tv_equal_half1"] = out["rof_equal"] = True
        for peer, sync in ((True, "signals"), (True, "barrier"), (False, "signals")):
            # halos read over NVLink inside the kernel (two ways of ordering the iterations) / sent as messages
            for half in (False, True):
                tv = ShardedPDTV(sh, (sh.nz_local, n, n), dev, half, peer_memory=peer, sync=sync, pairs=False)
                with single_iteration_tv():
                    whole = PD_TV_cupy(full, 4e-4, 9, 0, 1, 12.0, rank, half)
                for _ in range(2):  # buffers are reused across calls
                    part = tv(full[sh.z0:sh.z1].contiguous(), 4e-4, 9, 0, 1, 12.0)
                    out[f"tv_equal_half{int(half)}"] &= bool(torch.equal(sh.all_gather_volume(part), whole))
            if peer:
                # pairs of iterations per pass over peer memory (the default of ShardedPDTV for fp32 duals)
                tvp = ShardedPDTV(sh, (sh.nz_local, n, n), dev, False, peer_memory=True, sync=sync)
                assert tvp.pairs
                with single_iteration_tv():
                    whole = PD_TV_cupy(full, 4e-4, 9, 0, 1, 12.0, rank, False)
                for _ in range(2):
                    part = sh.all_gather_volume(tvp(full[sh.z0:sh.z1].contiguous(), 4e-4, 9, 0, 1, 12.0))
                    out["tv_pairs_maxdiff"] = max(out.get("tv_pairs_maxdiff", 0.0),
                                                  float((part - whole).abs().max() / whole.abs().max()))
            rof = ShardedROFTV(sh, (sh.nz_local, n, n), dev, False, peer_memory=peer, sync=sync)
            part = rof(full[sh.z0:sh.z1].contiguous(), 4e-4, 9, 1e-3)
            out["rof_equal"] &= bool(torch.equal(sh.all_gather_volume(part),
                                                 ROF_TV_cupy(full, 4e-4, 9, 1e-3, rank, False)))
        # --- sharded FISTA-OS + PD_TV == whole-volume run ------------------------------------------
        angles = np.linspace(0, np.pi, na, endpoint=False).astype(np.float32)
        sino = torch.rand((nz, na, n), generator=g).to(dev)
        alg = {"iterations": 3, "lipschitz_const": 2000.0, "nonnegativity": True, "recon_mask_radius": None}
        reg = {"method": "PD_TV", "regul_param": 3e-4, "iterations": 6}
        rec = RecToolsIRCuPy(n, 0, sh.nz_local, 0.0, angles, n, rank, 4)
        rec.set_zshard(sh)
        rec.tv_pairs = False  # one iteration per launch on both sides: bit for bit
        x_loc = rec.FISTA({"projection_data": sino[sh.z0:sh.z1].contiguous()}, dict(alg), dict(reg))
        x_all = sh.all_gather_volume(x_loc.contiguous())
        with single_iteration_tv():
            ref = RecToolsIRCuPy(n, 0, nz, 0.0, angles, n, rank, 4).FISTA({"projection_data": sino}, dict(alg),
                                                                          dict(reg))
        out["fista_equal"] = bool(torch.equal(x_all, ref))
        out["fista_maxdiff"] = float((x_all - ref).abs().max())
        # the defaults on both sides: pairs of iterations per pass, sharded (peer memory) and whole-volume
        rec.tv_pairs = None
        rec.set_zshard(sh)
        x_all = sh.all_gather_volume(rec.FISTA({"projection_data": sino[sh.z0:sh.z1].contiguous()}, dict(alg),
                                               dict(reg)).contiguous())
        ref = RecToolsIRCuPy(n, 0, nz, 0.0, angles, n, rank, 4).FISTA({"projection_data": sino}, dict(alg), dict(reg))
        out["fista_pairs_maxdiff"] = float((x_all - ref).abs().max() / ref.abs().max())
        # --- sharded ADMM-OS + ROF_TV == whole-volume run (BASELINE.json config 3 in miniature) ----------
        aalg = {"iterations": 3, "lipschitz_const": 2000.0, "ADMM_rho_const": 1.0, "ADMM_relax_par": 1.7,
                "recon_mask_radius": None}
        areg = {"method": "ROF_TV", "regul_param": 3e-4, "iterations": 6, "time_marching_step": 1e-3}
        a_loc = rec.ADMM({"projection_data": sino[sh.z0:sh.z1].contiguous()}, dict(aalg), dict(areg))
        a_ref = RecToolsIRCuPy(n, 0, nz, 0.0, angles, n, rank, 4).ADMM({"projection_data": sino}, dict(aalg), dict(areg))
        out["admm_equal"] = bool(torch.equal(sh.all_gather_volume(a_loc.contiguous()), a_ref))
        # --- sharded power method ~ whole-volume power method ---------------------------------------
        out["L_sharded"] = rec.powermethod({"projection_data": sino[sh.z0:sh.z1].contiguous()})
        out["L_whole"] = RecToolsIRCuPy(n, 0, nz, 0.0, angles, n, rank, 4).powermethod({"projection_data": sino})
        results[rank] = out
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_multi_gpu_sharded_tv_and_fista(world):
    """world = 4 has interior ranks (two neighbours each), world = 2 only boundary ranks."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    with tempfile.TemporaryDirectory() as d:
        mgr = mp.Manager()
        results = mgr.dict()
        mp.spawn(_worker, args=(world, os.path.join(d, "rdzv"), results), nprocs=world, join=True)
    for r in range(world):
        res = results[r]
        assert res["tv_equal_half0"] and res["tv_equal_half1"], res
        assert res["rof_equal"], res
        assert res["fista_equal"], res
        assert res["admm_equal"], res
        assert res["tv_pairs_maxdiff"] < 2e-6, res
        assert res["fista_pairs_maxdiff"] < 2e-6, res
        assert res["L_sharded"] == pytest.approx(res["L_whole"], rel=1e-3)
    assert all(results[r]["L_sharded"] == results[0]["L_sharded"] for r in range(world))
