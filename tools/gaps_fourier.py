"""GPU idle time inside one FOURIER_INV call at config 4: kernel timeline from torch.profiler (CUPTI), gaps between
consecutive kernels and memcpys on the device, largest first."""
import math
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tomobar_b200.methodsDIR_CuPy import RecToolsDIRCuPy  # noqa: E402

n, nz, na = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (2048, 128, 2000)
angles = np.linspace(0.0, math.radians(179.9), na).astype(np.float32)
R = RecToolsDIRCuPy(n, 0, nz, 0.0, angles, n, device_projector=0)
data = torch.rand((nz, na, n), device="cuda")
for _ in range(3):
    R.FOURIER_INV(data, filter_type="shepp", cutoff_freq=1.0)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    R.FOURIER_INV(data, filter_type="shepp", cutoff_freq=1.0)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
busy = sum(e.time_range.end - e.time_range.start for e in ev)
span = ev[-1].time_range.end - ev[0].time_range.start
print(f"{len(ev)} device activities, span {span / 1e3:.2f} ms, busy {busy / 1e3:.2f} ms, idle {(span - busy) / 1e3:.2f} ms")
gaps = []
for a, b in zip(ev[:-1], ev[1:]):
    g = b.time_range.start - a.time_range.end
    if g > 20:
        gaps.append((g, a.name[:60], b.name[:60]))
for g, a, b in sorted(gaps, reverse=True)[:25]:
    print(f"{g / 1e3:8.3f} ms  after {a}  before {b}")
by = {}
for e in ev:
    k = e.name[:60]
    by[k] = by.get(k, 0) + e.time_range.end - e.time_range.start
for k, v in sorted(by.items(), key=lambda kv: -kv[1])[:14]:
    print(f"{v / 1e3:8.2f} ms  {k}")
