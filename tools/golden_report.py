"""Runs every pinned golden of the reference's tests/test_RecToolsIRCuPy.py / test_RecToolsDIRCuPy.py
through the CUDA path and prints achieved vs pinned min / max (used to set the tolerances in
tests/test_gpu_goldens*.py).  python tools/golden_report.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_cases import CASES, run_case, load_scan  # noqa: E402


def main():
    scan = load_scan()
    for name, case in CASES.items():
        t0 = time.time()
        got = run_case(case, scan)
        torch.cuda.synchronize()
        line = [f"{name:34s}"]
        for key, want in case["expect"].items():
            g = got[key]
            line.append(f"{key}: got {g: .9g} want {want: .9g} rel {abs(g - want) / abs(want):.2e}")
        print("  ".join(line), f"[{time.time() - t0:.1f}s]", flush=True)


if __name__ == "__main__":
    main()
