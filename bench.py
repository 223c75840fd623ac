#!/usr/bin/env python
"""Benchmark of the hot path on B200: ordered-subsets FISTA + PD_TV on the headline geometry of
BASELINE.json (2048 x 2048 x 512 volume, 1800 angles) plus forward / back-projection rates.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A *step* is one ordered-subset sub-step of FISTA (subset forward projection with fused residual,
subset back-projection, gradient step, PD_TV prox, momentum).  `value` is outer FISTA iterations
per second (= sub-steps/s / OS) for the WHOLE volume; with N GPUs the volume is z-sharded
(512/N slices per rank, strong scaling).  The projector pair needs no communication; the 3-D TV
prox exchanges one-plane halos with the neighbouring shards between its inner iterations
(tomobar_b200.zshard.ShardedPDTV, bit-identical to the single-GPU prox; --independent-tv gives
the reference's HTTomo behaviour of independent z-blocks instead).  Prints ONE JSON line on rank 0.
"""

from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

HEADLINE = dict(n=2048, nz=512, na=1800, os=24, tv_iters=50, tv_lambda=3e-4)
METRIC = "fista_os_iterations_per_sec"
UNIT = "iter/s"
# BASELINE.json's configurations (SURVEY.md section 8, table of sizes and 8d for the algorithm parameters).  The
# default ("headline") is the size the metric is quoted on with the many-subset setting of the reference's demo
# (OS = 24: the TV prox is then 83 % of a sub-step, FP + BP 17 %; OS = 6 -- config 2's setting -- at this size gives
# 4x longer sub-steps with FP + BP at ~45 %).
CONFIGS = {
    "headline": dict(algo="fista", n=2048, nz=512, na=1800, os=24, tv_iters=50),
    "c1": dict(algo="fbp2d", n=256, nz=1, na=180, os=1, tv_iters=0),
    "c2": dict(algo="fista", n=1024, nz=256, na=900, os=6, tv_iters=50),
    "c3": dict(algo="admm", n=2048, nz=512, na=1800, os=24, tv_iters=30),
    "c4": dict(algo="fourier", n=2048, nz=128, na=2000, os=1, tv_iters=0),
    "c5": dict(algo="fista_ring", n=1536, nz=384, na=1500, os=6, tv_iters=50),
}
METRICS = {
    "fista": (METRIC, UNIT), "fista_ring": ("fista_os_huber_ring_iterations_per_sec", UNIT),
    "admm": ("admm_os_iterations_per_sec", UNIT), "fourier": ("fourier_inv_slices_per_sec", "slices/s"),
    "fbp2d": ("fbp2d_reconstructions_per_sec", "recon/s"),
}
C5_MODEL = dict(huber_threshold=0.05, ringGH_lambda=1e-4, ringGH_accelerate=50.0)


# ----------------------------------------------------------------------------------------------
# synthetic input: analytic ellipsoid phantom, closed-form parallel projections, Poisson noise
# (recipe of the reference's Demos/tomophantom_3D_recon1.py:24-70; TomoPhantom is not installed)
# ----------------------------------------------------------------------------------------------
ELLIPSOIDS = [
    # value, cx, cy, cz, ax, ay, az, phi   (unit cube coordinates in [-1, 1])
    (1.00, 0.00, 0.00, 0.00, 0.69, 0.92, 0.90, 0.0),
    (-0.80, 0.00, -0.0184, 0.00, 0.6624, 0.874, 0.88, 0.0),
    (-0.20, 0.22, 0.00, 0.00, 0.11, 0.31, 0.22, -0.31),
    (-0.20, -0.22, 0.00, 0.00, 0.16, 0.41, 0.28, 0.31),
    (0.10, 0.00, 0.35, -0.15, 0.21, 0.25, 0.41, 0.0),
    (0.10, 0.00, 0.10, 0.25, 0.046, 0.046, 0.05, 0.0),
    (0.10, -0.08, -0.605, 0.00, 0.046, 0.023, 0.05, 0.0),
    (0.10, 0.06, -0.605, 0.00, 0.023, 0.046, 0.02, 0.0),
]


def synth_sinogram(torch, nz_total, z0, z1, n, na, device, noise_seed=0, i0=8000.0):
    """Noisy post-log sinogram [z1-z0, na, n] of the ellipsoid phantom for slices z0..z1-1 of a
    volume with nz_total slices.  Line integrals are exact (closed form), in pixel units scaled so
    that the attenuation is O(1)."""
    angles = torch.linspace(0.0, math.radians(179.9), na, device=device, dtype=torch.float64)
    t = (torch.arange(n, device=device, dtype=torch.float64) - n / 2 + 0.5) / (n / 2)  # [-1, 1)
    zs = (torch.arange(z0, z1, device=device, dtype=torch.float64) - nz_total / 2 + 0.5) / (nz_total / 2)
    sino = torch.zeros((z1 - z0, na, n), device=device, dtype=torch.float32)
    for (val, cx, cy, cz, ax, ay, az, phi) in ELLIPSOIDS:
        hz = 1.0 - ((zs - cz) / az) ** 2                       # [z] cross-section scale^2
        inside = hz > 0
        sc = torch.sqrt(torch.clamp(hz, min=0.0))             # semi-axes shrink by sc
        th = angles - phi
        s2 = (ax * torch.cos(th)) ** 2 + (ay * torch.sin(th)) ** 2          # [a]
        tau = t[None, :] - (cx * torch.cos(angles) + cy * torch.sin(angles))[:, None]  # [a, u]
        # chord of the ellipse with semi-axes (ax*sc, ay*sc): 2 ax ay sc^2/s2' * sqrt(s2' - tau^2), s2' = s2 sc^2
        s2z = s2[None, :, None] * (sc ** 2)[:, None, None]
        chord = 2.0 * ax * ay * (sc ** 2)[:, None, None] / torch.clamp(s2z, min=1e-30) * torch.sqrt(
            torch.clamp(s2z - tau[None] ** 2, min=0.0))
        sino += (val * chord * inside[:, None, None]).to(torch.float32)
    sino *= 2.0  # attenuation scale: central chord ~ 2*0.2*... O(1)
    gen = torch.Generator(device=device).manual_seed(noise_seed + z0)
    counts = torch.poisson(i0 * torch.exp(-sino), generator=gen)
    sino = -torch.log(torch.clamp(counts, min=1.0) / i0)
    return sino.contiguous()


def synth_sinogram_numpy(nz_total, z0, z1, n, na):
    import torch

    return synth_sinogram(torch, nz_total, z0, z1, n, na, "cpu").numpy()


# ----------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            for k in ("hbm_gbs", "hbm_gb_s", "hbm"):
                if k in d:
                    return float(d[k]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle (port of the reference's loops + ASTRA's par3d / line-kernel
# models, C with OpenMP for the projector pair AND the TV operators) on ALL host cores, on a bounded sample,
# scaled linearly in slices x sub-steps ("extrapolated": true)
# ----------------------------------------------------------------------------------------------
def _oracle_all_cores():
    """The oracle with its OpenMP pool sized to the host: OMP_NUM_THREADS is set unconditionally (torchrun
    exports OMP_NUM_THREADS=1 to its workers) BEFORE liboracle.so and its libgomp are loaded.  Returns the
    module and the thread count the C code actually reports."""
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    from oracle import oracle as O

    O.build()
    return O, O.set_threads(cores)


def l2_policy(cfg, gpus):
    """Timing hygiene of the GPU arm, a function of the configuration only (both arms print the same config)."""
    if cfg["algo"] in ("fourier", "fbp2d"):
        return "L2 flushed (256 MB buffer rewritten) after every step"
    per_rank = 36.0 * (cfg["nz"] / gpus) * cfg["n"] * cfg["n"] / 1e9
    if per_rank > 1.0:
        return f"inputs larger than L2: one TV iteration streams {per_rank:.1f} GB per rank (L2 126 MB), no flush"
    return "L2 flushed (256 MB buffer rewritten) after every step"


def cpu_substep_rate(cfg, sample_slices=2):
    """One ordered-subset sub-step of the configuration's algorithm (or one direct reconstruction) on
    `sample_slices` slices through the oracle, all host cores; value in the configuration's metric."""
    O, threads = _oracle_all_cores()
    n, nz, na, os_n, algo = cfg["n"], cfg["nz"], cfg["na"], cfg["os"], cfg["algo"]
    angles = np.linspace(0.0, math.radians(179.9), na).astype(np.float32)
    if algo == "fbp2d":
        # BASELINE.json config 1: the reference's own CPU methodsDIR path, single-threaded like ASTRA's CPU BP,
        # and on all cores (angles dealt to threads)
        sino = synth_sinogram_numpy(1, 0, 1, n, na)[0]
        O.fbp2d_cpu(sino, angles, n)
        t1 = min(_timeit(lambda: O.fbp2d_cpu(sino, angles, n, threads=1)) for _ in range(3))
        tn = min(_timeit(lambda: O.fbp2d_cpu(sino, angles, n, threads=threads)) for _ in range(3))
        extra = {"single_thread_value": 1.0 / t1, "single_thread_ms": 1e3 * t1, "all_cores_ms": 1e3 * tn}
        return 1.0 / tn, threads, (f"whole workload, not a sample: _filtersinc2D + line-kernel BP (oracle/fbp2d_oracle.c) of "
                                   f"a {n}x{n} slice from {na} angles; {1e3 * t1:.1f} ms on one thread (ASTRA's CPU BP is "
                                   f"single-threaded), {1e3 * tn:.1f} ms on {threads} threads"), False, extra
    if algo == "fourier":
        # the reference has no CPU FOURIER_INV (methodsDIR.FOURIER is a 2-D scipy griddata toy): its CPU direct
        # method for 3-D data is FBP, which is what is timed beside the GPU's FOURIER_INV
        rec = O.RecDIR(n, 0, sample_slices, 0.0, angles, n)
        data = np.ascontiguousarray(np.swapaxes(synth_sinogram_numpy(nz, nz // 2, nz // 2 + sample_slices, n, na), 0, 1))
        t = _timeit(lambda: rec.FBP(data, cutoff_freq=1.0))
        return sample_slices / t, threads, (f"CPU direct method of the reference for 3-D data = FBP (sinc filter + "
                                            f"back-projection; it has no CPU FOURIER_INV) on {sample_slices} of {nz} "
                                            f"slices at N={n}, {na} angles; {t:.2f} s measured"), True, {}
    rec = O.RecIR(n, 0, sample_slices, 0.0, angles, n, os_n)
    b = synth_sinogram_numpy(nz, nz // 2, nz // 2 + sample_slices, n, na)
    x_t = np.zeros((sample_slices, n, n), np.float32)
    if algo == "admm":
        reg = {"method": "ROF_TV", "regul_param": cfg["tv_lambda"], "iterations": cfg["tv_iters"],
               "time_marching_step": 1e-3}
    else:
        reg = {"method": "PD_TV", "regul_param": cfg["tv_lambda"], "iterations": cfg["tv_iters"], "methodTV": 0,
               "PD_LipschitzConstant": 12.0}
    r_x = np.zeros((sample_slices, n), np.float32) if algo == "fista_ring" else None

    def substep():
        ind = rec._subset(0)
        if algo == "fista_ring":
            res, _ = rec.residual_ext(x_t, b[:, ind, :], True, 0, ind, None, "LS", C5_MODEL["huber_threshold"], r_x,
                                      C5_MODEL["ringGH_accelerate"], 0.1)
            g = rec._Atb(res, 0, True)
        else:
            g = rec.grad_data_term(x_t, b[:, ind, :], True, 0, ind, None, "LS")
        x = (x_t - np.float32(1e-4) * g).astype(np.float32)
        x = O.prox_regul(x, reg, 1)
        return x + np.float32(0.5) * (x - x_t)

    t = _timeit(substep)
    t_iter = t * (nz / sample_slices) * os_n  # one sub-step on the sample -> one outer iteration on the volume
    name = {"admm": "ROF_TV", "fista": "PD_TV", "fista_ring": "PD_TV"}[algo]
    return 1.0 / t_iter, threads, (f"one OS sub-step (FP+BP of {na // os_n} angles + {cfg['tv_iters']} {name} its, C/OpenMP) "
                                   f"on {sample_slices} of {nz} slices at N={n}, scaled x{nz // sample_slices} slices "
                                   f"x{os_n} subsets; {t:.2f} s measured"), True, {}


def _timeit(fn):
    t0 = time.perf_counter()
    fn()
    return time.perf_counter() - t0


def run_reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    metric, unit = METRICS[cfg["algo"]]
    vals = []
    for _ in range(max(1, min(args.steps, 3))):
        v, cores, sample, extrapolated, extra = cpu_substep_rate(cfg, sample_slices=2)
        vals.append(v)
    v = float(np.median(vals))
    steps_per_unit = cfg["os"] if cfg["algo"] in ("fista", "fista_ring", "admm") else (1.0 / cfg["nz"] if cfg["algo"] == "fourier" else 1)
    line = {
        "impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 / (v * steps_per_unit), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config(cfg, args.gpus),
        "cpu_baseline": dict({"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": sample,
                              "extrapolated": extrapolated}, **extra),
        "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "extrapolated": extrapolated,
    }
    print(json.dumps(line))


def _config(cfg, gpus):
    algo = cfg["algo"]
    vol = f"volume {cfg['n']}x{cfg['n']}x{cfg['nz']}, {cfg['na']} angles"
    if algo == "admm":
        workload = f"ADMM-OS + ROF_TV, {vol}, OS={cfg['os']}, rho 1, alpha 1.7, ROF_TV {cfg['tv_iters']} inner iterations"
    elif algo == "fourier":
        workload = f"FOURIER_INV (USFFT gridding, filter shepp, cutoff 1.0), {vol}"
    elif algo == "fbp2d":
        workload = f"2-D FBP (sinc filter a = 1.1 + back-projection), slice {cfg['n']}x{cfg['n']}, {cfg['na']} angles"
    elif algo == "fista_ring":
        workload = (f"FISTA-OS, Huber data term (threshold {C5_MODEL['huber_threshold']}) + Group-Huber ring model "
                    f"(lambda {C5_MODEL['ringGH_lambda']}, accelerate {C5_MODEL['ringGH_accelerate']:g}) + PD_TV, {vol}, "
                    f"OS={cfg['os']}, PD_TV {cfg['tv_iters']} inner iterations; robust terms are an extension whose parity is "
                    "UNPINNED (no code in the reference snapshot, SURVEY.md 8a row H)")
    else:
        workload = (f"FISTA-OS + PD_TV, {vol}, OS={cfg['os']}, PD_TV {cfg['tv_iters']} inner iterations (fp32 duals)")
    return {
        "name": cfg.get("name", "custom"), "workload": workload,
        "n": cfg["n"], "nz": cfg["nz"], "angles": cfg["na"], "os_number": cfg["os"],
        "tv_inner_iterations": cfg["tv_iters"], "z_shards": gpus,
        "tv_across_shards": cfg.get("halo", "n/a") if gpus > 1 and cfg["tv_iters"] else "n/a",
        "l2_policy": l2_policy(cfg, gpus),
    }


def traffic_of(kernel, voxels, half=False):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this
    very launch shape) from the committed profiles/ncu_traffic_r0*.json, newest round first."""
    for name in ("ncu_traffic_r02.json", "ncu_traffic_r01.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as fh:
                for rec_t in json.load(fh):
                    if (rec_t["kernel"] == kernel and rec_t["voxels"] == voxels
                            and bool(rec_t.get("half", False)) == bool(half)):
                        return float(rec_t["dram_bytes_per_launch"])
        except (OSError, ValueError, KeyError):
            continue
    return None


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="tmb", choices=["tmb", "reference"])
    ap.add_argument("--timeline", default=None, metavar="FILE",
                    help="after the timed region, run one more step under torch.profiler and write the device timeline "
                         "(busy / idle time, largest gaps, time per kernel) of this rank to FILE.rank<r> (iterative configs)")
    ap.add_argument("--config", default="headline", choices=sorted(CONFIGS),
                    help="BASELINE.json configuration: headline (default: FISTA-OS 24 + PD_TV at 2048^2 x 512 / 1800), "
                         "c1 (2-D FBP 256^2, the CPU methodsDIR case), c2 (FISTA-OS 6 + PD_TV, 1024^2 x 256 / 900), "
                         "c3 (ADMM-OS 24 + ROF_TV, 2048^2 x 512 / 1800), c4 (FOURIER_INV 2048^2 x 128 / 2000), "
                         "c5 (FISTA Huber + ring model, 1536^2 x 384 / 1500)")
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--nz", type=int, default=None)
    ap.add_argument("--angles", type=int, default=None)
    ap.add_argument("--os", type=int, default=None)
    ap.add_argument("--tv-iters", type=int, default=None)
    ap.add_argument("--half", action="store_true", help="fp16 storage of the TV dual variables")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--algo", default=None, choices=["fista", "admm"],
                    help="(kept from round 1) --algo admm == --config c3")
    ap.add_argument("--halo-messages", action="store_true",
                    help="multi-GPU: refresh the TV ghost planes with NCCL send/recv instead of letting the "
                         "kernel read the neighbours' planes over NVLink (peer memory)")
    ap.add_argument("--halo-barrier", action="store_true",
                    help="multi-GPU peer-memory halos: one all-rank barrier per TV iteration instead of "
                         "pairwise semaphores with the two neighbours")
    ap.add_argument("--tv-single", action="store_true",
                    help="multi-GPU peer-memory halos: one PD_TV iteration per launch (round 1's behaviour) instead of "
                         "pairs of iterations per pass (tmb_pd_tv_iter2, one neighbour synchronisation per pair)")
    ap.add_argument("--independent-tv", action="store_true",
                    help="multi-GPU: TV per z-shard without halo exchange (seams at the shard borders)")
    args = ap.parse_args()
    name = "c3" if args.algo == "admm" and args.config == "headline" else args.config
    cfg = dict(CONFIGS[name], name=name, tv_lambda=HEADLINE["tv_lambda"])
    for key, val in (("n", args.n), ("nz", args.nz), ("na", args.angles), ("os", args.os), ("tv_iters", args.tv_iters)):
        if val is not None:
            cfg[key] = val
            cfg["name"] = name + " (resized)"
    cfg["halo"] = ("independent z-blocks" if args.independent_tv else
                   "exact, NCCL messages between inner iterations" if args.halo_messages else
                   "exact, peer loads over NVLink inside the TV kernel"
                   + (", all-rank barrier" if args.halo_barrier else ", pairwise semaphores")
                   + (" per iteration" if (args.tv_single or args.half or cfg["algo"] == "admm") else
                      " per PAIR of iterations (two iterations per pass)"))
    if args.warmup < 3:
        args.warmup = 3

    if args.impl == "reference":
        run_reference_arm(args, cfg)
        return
    if cfg["algo"] in ("fourier", "fbp2d"):
        run_direct(args, cfg)
        return
    run_iterative(args, cfg)


def _setup_device():
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: tomobar_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    return torch, dist, world, rank, local_rank, dev


def _timed(torch, fn, reps):
    fn()
    torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    c.record()
    torch.cuda.synchronize()
    return a.elapsed_time(c) / reps


def run_direct(args, cfg):
    """Configs 4 (FOURIER_INV) and 1 (2-D FBP): a step is one whole reconstruction; slices are independent, so N
    ranks reconstruct N z-blocks (c4) or N replicas (c1: a single slice does not shard)."""
    torch, dist, world, rank, local_rank, dev = _setup_device()
    from tomobar_b200._lib import lib, check
    from tomobar_b200._tensors import ptr
    from tomobar_b200.methodsDIR_CuPy import RecToolsDIRCuPy
    from tomobar_b200.zshard import ZShard

    n, nz, na, algo = cfg["n"], cfg["nz"], cfg["na"], cfg["algo"]
    metric, unit = METRICS[algo]
    fourier = algo == "fourier"
    shard = ZShard(nz) if fourier else None
    z0, z1 = (shard.z0, shard.z1) if fourier else (0, 1)
    nz_loc = z1 - z0
    angles = np.linspace(0.0, math.radians(179.9), na).astype(np.float32)
    rec = RecToolsDIRCuPy(n, 0, nz_loc, 0.0, angles, n, device_projector=local_rank)
    b = torch.empty((nz_loc, na, n), dtype=torch.float32, device=dev)
    for s in range(0, nz_loc, 16):
        e = min(nz_loc, s + 16)
        b[s:e] = synth_sinogram(torch, nz, z0 + s, z0 + e, n, na, dev)
    b_fbp = b.swapaxes(0, 1).contiguous() if not fourier else None  # FBP's default axes: [angles, detY, detX]
    torch.cuda.synchronize()

    def step():
        if fourier:
            return rec.FOURIER_INV(b, filter_type="shepp", cutoff_freq=1.0, recon_mask_radius=None)
        return rec.FBP(b_fbp, cutoff_freq=1.1, recon_mask_radius=None)

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # the L2 flush sits between the timed steps, outside the event pairs
    ms_total = 0.0
    barrier()
    for _ in range(args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush_buf.zero_()
        e0.record()
        step()
        e1.record()
        torch.cuda.synchronize()
        ms_total += e0.elapsed_time(e1)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tt = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    ms_step = ms_total / args.steps
    units = float(nz) if fourier else float(world)  # slices of the whole volume / one slice per replica
    value = units / (ms_step * 1e-3)

    st = torch.cuda.current_stream(dev).cuda_stream
    peak, peak_src = measured_peak_hbm()
    if fourier:
        # dominant kernel: the polar -> Cartesian gather (tmb_fi_gather), timed alone on buffers of the sizes the step
        # launches it on: FOURIER_INV grids, transforms and unpads chunk by chunk of complex slices
        nz2 = nz_loc // 2
        chunk = max(1, min(nz2, (1 << 28) // (4 * n * n)))
        n_chunks = -(-nz2 // chunk)
        theta = torch.as_tensor(-angles, dtype=torch.float32, device=dev)
        sorted_theta, sorted_idx = torch.sort(theta)
        sorted_idx = sorted_idx.to(torch.int32)
        datac = torch.randn((chunk, na, n), dtype=torch.complex64, device=dev)
        fde = torch.empty((chunk, 2 * n, 2 * n), dtype=torch.complex64, device=dev)
        mu = -np.log(1e-4) / (2 * n * n)
        m = int(np.ceil(2 * n * 1 / np.pi * np.sqrt(-mu * np.log(1e-4) + (mu * n) * (mu * n) / 4)))
        # (the entry point FOURIER_INV calls for these sizes: polar samples stored as slice pairs when every chunk is
        # whole blocks of 8 complex slices)
        gather = lib.tmb_fi_gather_pairs if (nz2 % 8 == 0 and chunk % 8 == 0) else lib.tmb_fi_gather
        ms_k = _timed(torch, lambda: check(gather(ptr(datac), ptr(fde), ptr(theta), ptr(sorted_theta),
                                                  ptr(sorted_idx), m, float(np.float32(mu)), n, na, chunk, st),
                                           "tmb_fi_gather"), 6)
        bytes_k = 8.0 * chunk * na * n + 8.0 * chunk * 4 * n * n  # polar samples read once, grid written once
        kname = (f"k_fi_gather_w (USFFT gather onto the 2n x 2n grid, {chunk} complex slices per launch, {n_chunks} launches "
                 "per step; a warp walks the polar lines of its patch in lock step, samples read as slice pairs: bound by the "
                 "L1 data pipe (ncu: LSU wavefronts 72 % of peak), the grid write is its algorithmic HBM traffic)")
        traffic = traffic_of("k_fi_gather_w", int(chunk) * 4 * n * n)
        del datac, fde
        # pad + crop per filter chunk, scale-sign, gather + unpad per grid chunk (+ torch's spectrum product and cuFFT's
        # own kernels, not counted)
        per = max(2, ((1 << 27) // (na * 2 ** int(np.ceil(np.log2(3 * n))))) // 2 * 2)  # slices per filter chunk
        launches = 2 * -(-nz_loc // per) + 1 + 2 * n_chunks
        ms_k_step = ms_k * n_chunks
    else:
        sino_f = torch.randn((nz_loc, na, n), device=dev)
        vol = torch.empty((nz_loc, n, n), device=dev)
        A = rec.Atools
        ms_k = _timed(torch, lambda: check(lib.tmb_bp3d(A._g, -1, ptr(sino_f), ptr(vol), ptr(A._workspace()), st), "bp"), 5)
        bytes_k = 4.0 * (nz_loc * n * n + nz_loc * na * n)
        kname = "k_bp (voxel-driven back-projection: shared-memory-bandwidth bound, HBM fraction tiny by construction)"
        traffic = None
        launches = 4
        ms_k_step = ms_k
    roofline = {"kernel": kname, "bound": "hbm", "achieved": bytes_k / (ms_k * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": bytes_k / (ms_k * 1e-3) / 1e9 / peak, "peak_source": peak_src, "traffic": traffic,
                "algorithmic_bytes_per_launch": bytes_k, "ms_per_launch": ms_k, "share_of_step": ms_k_step / ms_step}

    e2e = None
    if not args.no_e2e:
        src = b if fourier else b_fbp
        b_host = torch.empty(src.shape, dtype=torch.float32, pin_memory=True)
        b_host.copy_(src)
        out_shape = (nz, n, n) if (fourier and rank == 0) else (nz_loc, n, n)
        out_host = torch.empty(out_shape, dtype=torch.float32, pin_memory=True)

        def e2e_iter():
            nonlocal b, b_fbp
            d = b_host.to(dev, non_blocking=True)
            if fourier:
                b = d
            else:
                b_fbp = d
            r = step()
            if fourier and world > 1:
                r = shard.all_gather_volume(r.contiguous())  # the one collective: final volume assembly
                if rank != 0:
                    return
            out_host.copy_(r, non_blocking=True)

        e2e_iter()
        barrier()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            e2e_iter()
        barrier()
        dt = (time.perf_counter() - t0) / reps
        if world > 1:
            tt = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        e2e = {"value": units / dt, "unit": unit, "h2d_bytes_per_step": int(b_host.numel() * 4 * (world if fourier else 1)),
               "d2h_bytes_per_step": int(nz * n * n * 4) if fourier else int(out_host.numel() * 4),
               "note": f"RecToolsDIRCuPy.{'FOURIER_INV' if fourier else 'FBP'} from a pinned host sinogram to a pinned host "
                       "volume" + ("; z-blocks all-gathered, rank 0 copies the whole volume back" if world > 1 and fourier else "")}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample, extrapolated, extra = cpu_substep_rate(cfg, sample_slices=2)
        cpu_baseline = dict({"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": sample,
                             "extrapolated": extrapolated}, **extra)
    if rank == 0:
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong" if fourier else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": _config(cfg, world), "clocks": clocks, "e2e": e2e, "gpu_launches": launches * args.steps,
                "roofline": roofline, "cpu_baseline": cpu_baseline}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_iterative(args, cfg):
    torch, dist, world, rank, local_rank, dev = _setup_device()
    from tomobar_b200._lib import lib, check
    from tomobar_b200._tensors import ptr
    from tomobar_b200.methodsIR_CuPy import RecToolsIRCuPy
    from tomobar_b200.regularisersCuPy import PD_TV_cupy
    from tomobar_b200.zshard import ZShard

    metric, unit = METRICS[cfg["algo"]]
    n, nz, na, os_n = cfg["n"], cfg["nz"], cfg["na"], cfg["os"]
    # z-shard: contiguous block of slices per rank
    shard = ZShard(nz)
    z0, z1, nz_loc = shard.z0, shard.z1, shard.nz_local
    angles = np.linspace(0.0, math.radians(179.9), na).astype(np.float32)

    rec = RecToolsIRCuPy(n, 0, nz_loc, 0.0, angles, n, local_rank, os_n)
    if world > 1 and not args.independent_tv:
        rec.set_zshard(shard)
        rec.tv_peer_memory = False if args.halo_messages else None
        rec.tv_sync = "barrier" if args.halo_barrier else "signals"
        rec.tv_pairs = False if args.tv_single else None
    rec.nonneg_regul = 1
    A = rec.Atools
    # synthetic data generated on the device, slice blocks of 16 to bound temporaries
    b = torch.empty((nz_loc, na, n), dtype=torch.float32, device=dev)
    for s in range(0, nz_loc, 16):
        e = min(nz_loc, s + 16)
        b[s:e] = synth_sinogram(torch, nz, z0 + s, z0 + e, n, na, dev)
    torch.cuda.synchronize()

    reg = {"method": "PD_TV", "regul_param": cfg["tv_lambda"], "iterations": cfg["tv_iters"], "methodTV": 0,
           "PD_LipschitzConstant": 12.0, "half_precision": bool(args.half)}
    st = torch.cuda.current_stream(dev).cuda_stream
    vol_shape = A.vol_geom
    count = nz_loc * n * n
    X = torch.zeros(vol_shape, device=dev)
    X_old = torch.zeros(vol_shape, device=dev)
    X_t = torch.zeros(vol_shape, device=dev)
    G = torch.empty(vol_shape, device=dev)
    L_inv = 1.0 / 2.0e4  # fixed step: the benchmark times the loop, not the power method
    state = {"t": np.float32(1.0), "sub": 0}

    admm = cfg["algo"] == "admm"
    ring = cfg["algo"] == "fista_ring"
    if admm:
        # ADMM state (methodsIR_CuPy.py:531-566): x, z, z_old, u; G doubles as the prox input
        reg = {"method": "ROF_TV", "regul_param": cfg["tv_lambda"], "iterations": cfg["tv_iters"],
               "time_marching_step": 1e-3, "half_precision": bool(args.half)}
        Zv, Zo, Uv = torch.zeros(vol_shape, device=dev), torch.zeros(vol_shape, device=dev), X_old
        tau_admm, rho = 0.9 / (2.0e4 + 1.0), 1.0
    if ring:
        # Group-Huber ring model (methodsIR_CuPy.FISTA): one offset per detector pixel with its own momentum
        rs = {"r": torch.zeros((nz_loc, A.nu), device=dev), "r_x": torch.zeros((nz_loc, A.nu), device=dev),
              "vec": torch.empty((nz_loc, A.nu), device=dev)}

    def substep_admm():
        A.grad_data_term(Zv, b, state["sub"], "LS", None, out=G)
        check(lib.tmb_admm_z_step(ptr(Zv), ptr(Zo), ptr(X), ptr(Uv), ptr(G), ptr(X_t), count, tau_admm, rho, 1, 1,
                                  1.7, st), "admm_z")
        rec._prox_into(X_t, reg, X)
        state["sub"] = (state["sub"] + 1) % os_n
        if state["sub"] == 0:
            check(lib.tmb_admm_u_step(ptr(Uv), ptr(Zv), ptr(X), count, st), "admm_u")

    def substep():
        nonlocal X, X_old
        if admm:
            return substep_admm()
        X_old, X = X, X_old
        t_old = state["t"]
        if ring:
            A.grad_data_term_ext(X_t, b, state["sub"], "LS", None, C5_MODEL["huber_threshold"], rs["r_x"],
                                 C5_MODEL["ringGH_accelerate"], 0.1, rs["vec"], out=G)
            r_old = rs["r"]
            rs["r"] = rs["r_x"] - np.float32(L_inv) * rs["vec"]
        else:
            A.grad_data_term(X_t, b, state["sub"], "LS", None, out=G)
        check(lib.tmb_fista_grad_step(ptr(X_t), ptr(G), ptr(G), count, L_inv, 1, st), "grad_step")
        rec._prox_into(G, reg, X)  # PD_TV; whole-volume across the shards when world > 1
        t = np.float32((1.0 + np.sqrt(1.0 + 4.0 * t_old ** 2)) * 0.5)
        coef = float((t_old - 1.0) / t)
        check(lib.tmb_fista_momentum(ptr(X), ptr(X_old), ptr(X_t), count, coef, st), "momentum")
        if ring:
            r = torch.clamp(rs["r"].abs() - np.float32(C5_MODEL["ringGH_lambda"]), min=0) * torch.sign(rs["r"])
            rs["r"], rs["r_x"] = r, (r + coef * (r - r_old)).contiguous()
        state["t"] = t
        state["sub"] = (state["sub"] + 1) % os_n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # L2 hygiene: one TV iteration streams 36 B/voxel; when that is far above the 126 MB L2 nothing
    # survives between iterations, otherwise a buffer larger than the L2 is rewritten after every step
    stream_gb = 36.0 * count / 1e9
    flush_buf = None if stream_gb > 1.0 else torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    # our kernels per sub-step: layout conversion, forward projector (k_fpq [+ k_fp_finish] per chunk),
    # k_bp, gradient / z step, TV iterations, momentum (ADMM: the u update once per outer iteration)
    fp_launches = max(1, lib.tmb_geom_fp_launches(A._g, 0))
    # PD_TV: pairs of iterations per launch where the fused kernel applies -- the unsharded prox (tmb_pd_tv) and,
    # over peer memory, the z-sharded one (tmb_pd_tv_iter2); message halos and fp16 duals launch every iteration
    pairs_sharded = world > 1 and not (args.tv_single or args.halo_messages or args.half or args.independent_tv)
    if admm:
        tv_launches = cfg["tv_iters"]
    elif world > 1 and not args.independent_tv:
        tv_launches = (cfg["tv_iters"] + 1) // 2 if pairs_sharded else cfg["tv_iters"]
    else:
        tv_launches = lib.tmb_pd_tv_launches(nz_loc, n, n, cfg["tv_iters"], int(bool(args.half)))
    launches_per_step = 1 + fp_launches + 1 + 1 + tv_launches + 1 + (1 if ring else 0)

    for _ in range(args.warmup):
        substep()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        substep()
        if flush_buf is not None:
            flush_buf.zero_()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tt = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    ms_step = ms_total / args.steps
    value = 1000.0 / (ms_step * os_n)  # outer iterations per second for the whole (sharded) volume

    if args.timeline:
        # outside the timed region: one more sub-step under the profiler (device busy / idle time, gaps, kernels)
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from timeline import device_timeline

        barrier()
        report = device_timeline(substep, f"{cfg.get('name', args.config)} sub-step, rank {rank} of {world} "
                                          f"({ms_step:.2f} ms per step in the timed region)")
        with open(f"{args.timeline}.rank{rank}", "w") as fh:
            fh.write(report + "\n")
        barrier()

    # ---- per-kernel timings (CUDA events on the launching stream) ------------------------------
    def timed(fn, reps):
        return _timed(torch, fn, reps)

    na_s = A.subset_size(0)
    sub_sino = torch.empty((nz_loc, na_s, n), device=dev)
    ms_fp = timed(lambda: check(lib.tmb_fp3d(A._g, 0, ptr(X_t), ptr(sub_sino), ptr(A._workspace()), st), "fp"), 3)
    ms_bp = timed(lambda: check(lib.tmb_bp3d(A._g, 0, ptr(sub_sino), ptr(G), ptr(A._workspace()), st), "bp"), 3)
    tv_reps = max(4, cfg["tv_iters"])
    if admm:
        from tomobar_b200.regularisersCuPy import ROF_TV_cupy

        ms_tv = timed(lambda: ROF_TV_cupy(G, reg["regul_param"], tv_reps, 1e-3, local_rank, reg["half_precision"],
                                          out=X), 2) / tv_reps
    else:
        ms_tv = timed(lambda: PD_TV_cupy(G, reg["regul_param"], tv_reps, 0, 1, 12.0, local_rank,
                                         reg["half_precision"], out=X), 2) / tv_reps
    upd_sub = float(nz_loc) * n * n * na_s
    # algorithmic bytes of ONE launch: every array read once and written once (in, U, P1..P3 in; U, P1..P3 out)
    bytes_tv = (12.0 if admm else (24.0 if args.half else 36.0)) * count
    peak, peak_src = measured_peak_hbm()
    # the kernel-only timing above goes through the unsharded entry point (for PD_TV: memset / copy-free first pass,
    # then pairs -- the same kernel family the z-sharded default launches)
    tv_rep_launches = tv_reps if admm else lib.tmb_pd_tv_launches(nz_loc, n, n, tv_reps, int(bool(args.half)))
    fused_tv = (not admm) and tv_rep_launches < tv_reps
    ms_tv_launch = ms_tv * tv_reps / tv_rep_launches
    tv_gbs = bytes_tv / (ms_tv_launch * 1e-3) / 1e9
    if world > 1 and fused_tv and not pairs_sharded:
        # this sharded step launches single iterations (strip kernel): time that kernel for the roofline
        old_mode = lib.tmb_tv_set_simple_kernels(3)
        try:
            ms_tv = timed(lambda: PD_TV_cupy(G, reg["regul_param"], tv_reps, 0, 1, 12.0, local_rank,
                                             reg["half_precision"], out=X), 2) / tv_reps
        finally:
            lib.tmb_tv_set_simple_kernels(old_mode)
        fused_tv, ms_tv_launch = False, ms_tv
        tv_gbs = bytes_tv / (ms_tv_launch * 1e-3) / 1e9
    share_tv = ms_tv * cfg["tv_iters"] / ms_step
    kernel_id = "k_rof_tv3d_w" if admm else ("k_pd_tv3d_f2s" if fused_tv else "k_pd_tv3d_w")
    traffic = traffic_of(kernel_id, count, args.half)
    roofline = {
        "kernel": ("k_rof_tv3d_w (one fused ROF iteration; instruction-bound, 12 B/voxel)" if admm else
                   ("k_pd_tv3d_f2s (TWO Chambolle-Pock iterations per launch, nothing stored in between: 18 B/voxel "
                    "per iteration; co-limited by the LSU pipe, instruction issue and HBM, DESIGN.md 4.3b)" if fused_tv else
                    "k_pd_tv3d_w (one Chambolle-Pock iteration, warp-strip kernel)")),
        "bound": "hbm", "achieved": tv_gbs, "peak": peak,
        "unit": "GB/s", "frac": tv_gbs / peak, "peak_source": peak_src, "traffic": traffic,
        "algorithmic_bytes_per_launch": bytes_tv, "ms_per_launch": ms_tv_launch,
        "iterations_per_launch": 2 if fused_tv else 1, "share_of_step": share_tv,
        # SURVEY.md 8(d) counts PD_TV at 36 B/voxel per ITERATION (the reference's one-launch-per-iteration
        # structure); against that figure a two-iteration launch scores above the copy peak
        "per_iteration_equivalent": {"gbs": bytes_tv / (ms_tv * 1e-3) / 1e9, "frac": bytes_tv / (ms_tv * 1e-3) / 1e9 / peak,
                                     "ms_per_iteration": ms_tv},
    }
    kernels = {
        "fp_subset_ms": ms_fp, "bp_subset_ms": ms_bp, "tv_iteration_ms": ms_tv,
        "fp_gups": upd_sub / (ms_fp * 1e-3) / 1e9 * world, "bp_gups": upd_sub / (ms_bp * 1e-3) / 1e9 * world,
        "fp_gproj_per_s": float(nz_loc) * na_s * n / (ms_fp * 1e-3) / 1e9 * world,
        "bp_gproj_per_s": float(nz_loc) * na_s * n / (ms_bp * 1e-3) / 1e9 * world,
        "lds_roof_gups_per_gpu": 148 * 16 * 1.965e9 / 1e9,
    }

    # ---- end to end through the public class with HOST buffers -----------------------------------
    e2e = None
    if not args.no_e2e:
        del sub_sino
        b_host = torch.empty(b.shape, dtype=torch.float32, pin_memory=True)
        b_host.copy_(b)
        # N > 1: the final volume is assembled with the one all-gather the path has and rank 0 returns it to the host
        out_host = torch.empty((nz, n, n) if (world == 1 or rank == 0) else (1,), dtype=torch.float32, pin_memory=True)
        del X, X_old, X_t, G
        if admm:
            del Zv, Zo
        torch.cuda.empty_cache()
        data_model = {}
        if ring:
            data_model = {"huber_threshold": C5_MODEL["huber_threshold"], "ringGH_lambda": C5_MODEL["ringGH_lambda"],
                          "ringGH_accelerate": C5_MODEL["ringGH_accelerate"]}

        def e2e_iter():
            d = b_host.to(dev, non_blocking=True)
            alg = {"iterations": 1, "lipschitz_const": 2.0e4, "nonnegativity": True, "recon_mask_radius": None}
            if admm:
                alg.update({"ADMM_rho_const": 1.0, "ADMM_relax_par": 1.7})
            r = (rec.ADMM if admm else rec.FISTA)(dict({"projection_data": d}, **data_model), alg, dict(reg))
            if world > 1:
                r = shard.all_gather_volume(r.contiguous())
                if rank != 0:
                    return
            out_host.copy_(r, non_blocking=True)

        e2e_iter()
        barrier()
        t0 = time.perf_counter()
        reps = 1
        for _ in range(reps):
            e2e_iter()
        barrier()
        dt = (time.perf_counter() - t0) / reps
        if world > 1:
            tt = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        e2e = {"value": 1.0 / dt, "unit": unit, "h2d_bytes_per_step": int(nz * na * n * 4 / os_n),
               "d2h_bytes_per_step": int(nz * n * n * 4 / os_n),
               "note": f"RecToolsIRCuPy.{'ADMM' if admm else 'FISTA'}(iterations=1) from pinned host sinogram(s) to a "
                       "pinned host volume" + ("; z-shards all-gathered (the path's one collective), rank 0 copies the "
                                               "whole volume back" if world > 1 else "")}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample, extrapolated, extra = cpu_substep_rate(cfg, sample_slices=2)
        cpu_baseline = dict({"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": sample,
                             "extrapolated": extrapolated}, **extra)

    if rank == 0:
        line = {
            "metric": metric, "value": value, "unit": unit,
            "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": _config(cfg, world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            # BASELINE.json's second metric, whole job: line integrals per second of one subset projection
            "fp_gproj_per_s": kernels["fp_gproj_per_s"], "bp_gproj_per_s": kernels["bp_gproj_per_s"],
            "fp_gups": kernels["fp_gups"], "bp_gups": kernels["bp_gups"], "kernels": kernels,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
