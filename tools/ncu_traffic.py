"""Turns an ncu report of the TV / gather kernels into profiles/ncu_traffic_r02.json (DRAM bytes per launch):
   python tools/ncu_traffic.py gpurun_out/prof_traffic.ncu-rep nz n      (k_fi_gather: nz = complex slices, n = 2 * width)"""
import csv
import io
import json
import os
import subprocess
import sys

rep, nz, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
out = []
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    kern = next((k for k in ("k_pd_tv3d_f2s", "k_pd_tv3d_f2", "k_pd_tv3d_w", "k_rof_tv3d_w", "k_fi_gather_w", "k_fi_gather") if k in name), None)
    if kern is None:
        continue
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[idx[m]]) * scale[units[idx[m]]]
    out.append({"kernel": kern, "half": "__half" in name, "voxels": nz * n * n, "dram_bytes_per_launch": tot,
                "duration_ms_under_ncu": float(r[idx["gpu__time_duration.sum"]]) *
                {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[units[idx["gpu__time_duration.sum"]]],
                "source": os.path.basename(rep),
                **{m: float(r[idx[m]]) for m in ("smsp__issue_active.avg.pct_of_peak_sustained_active",
                                                 "sm__warps_active.avg.pct_of_peak_sustained_active",
                                                 "lts__t_sector_hit_rate.pct") if m in idx}})
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic_r02.json")
old = []
if os.path.exists(path):
    old = [o for o in json.load(open(path)) if (o["kernel"], o["half"], o["voxels"]) not in
           {(x["kernel"], x["half"], x["voxels"]) for x in out}]
json.dump(old + out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
