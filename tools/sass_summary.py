"""SASS summary of the built library (cuobjdump -sass): architectures, opcode counts of the families that matter for the
design (bulk copies, barriers, tensor-core / TMA-tensor opcodes that must be absent, memory, arithmetic), entry points.
python tools/sass_summary.py > profiles/sass_summary_r02.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "tomobar_b200", "libtmb.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
archs = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
funcs = re.findall(r"Function : (\S+)", txt)
ops = collections.Counter()
for line in txt.splitlines():
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m:
        ops[m.group(1)] += 1
print("SASS summary of tomobar_b200/libtmb.so (cuobjdump -sass, end of round 2; tools/sass_summary.py)")
print(f"architectures: {archs}   kernels: {len(funcs)}   instructions: {sum(ops.values())}")
for k in ("UBLKCP", "UBLKPF", "SYNCS", "UTMALDG", "UTMASTG", "UTCMMA", "HMMA", "LDG", "STG", "LDS", "STS", "LDGSTS", "SHFL",
          "MUFU", "FFMA", "FFMA2", "FADD", "FMUL", "ATOMG", "REDG", "CCTL", "BAR"):
    print(f"  {k:8s} {ops.get(k, 0):8d}")
print("top opcodes: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(25)))
names = sorted(set(subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()))
short = sorted(set(re.sub(r"<.*", "", n.split("(")[0]) for n in names))
print("kernel entry points (template arguments dropped where they only multiply variants):")
for n in short:
    print("  " + n)
