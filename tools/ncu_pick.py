"""Print selected metrics of an `ncu --page raw --csv` export: python tools/ncu_pick.py file.csv [substring ...]"""
import csv
import sys

DEFAULT = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "lts__t_bytes.sum",
           "smsp__inst_executed.sum", "local_op", "op_local", "op_shared", "op_global", "lts__t_sector_hit_rate.pct",
           "sm__warps_active.avg.pct", "issue_stalled", "launch__registers", "smsp__issue_active.avg.pct",
           "l1tex__data_pipe_lsu_wavefronts", "sm__inst_executed_pipe"]
rows = list(csv.reader(open(sys.argv[1])))
pats = sys.argv[2:] or DEFAULT
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:80] if "Kernel Name" in hdr else "")
    for i, h in enumerate(hdr):
        if any(p in h for p in pats) and "not_issued" not in h and r[i] not in ("", "0", "n/a"):
            print(f"  {h:90s} {units[i]:14s} {r[i]}")
