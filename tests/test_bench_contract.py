"""bench.py's reference arm (the CPU port of the hot path) runs without a GPU and prints the JSON
line of the bench contract."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the arm must size its own pool (VERDICT r1, weak point 9)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "48", "--nz", "4",
                          "--angles", "24", "--os", "4", "--tv-iters", "2", "--steps", "1", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "fista_os_iterations_per_sec"
    assert line["value"] > 0 and line["unit"] == "iter/s" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert line["extrapolated"] is True and line["cpu_baseline"]["extrapolated"] is True
    assert line["e2e"] == {"value": line["value"], "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["n"] == 48 and line["config"]["os_number"] == 4


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_arm_config1_is_the_cpu_fbp_path():
    """BASELINE.json config 1: the reference's CPU methodsDIR FBP, single-threaded and on all cores, core count stated."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c1", "--steps", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    cb = line["cpu_baseline"]
    assert line["metric"] == "fbp2d_reconstructions_per_sec" and line["config"]["n"] == 256 and line["config"]["angles"] == 180
    assert cb["single_thread_value"] > 0 and cb["value"] > 0 and cb["cores"] >= 1 and line["extrapolated"] is False


def test_both_arms_describe_the_same_config():
    """`config` is a function of the configuration and the GPU count only, so the driver's same_config check holds."""
    sys.path.insert(0, ROOT)
    import bench

    for name, c in bench.CONFIGS.items():
        cfg = dict(c, name=name, tv_lambda=3e-4, halo="x")
        assert bench._config(cfg, 1) == bench._config(dict(cfg), 1)
        assert "n/a" not in bench._config(cfg, 1)["l2_policy"]
