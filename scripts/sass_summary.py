"""Opcode counts per kernel of libcmpt_b200.so (cuobjdump -sass): which kernels carry TMA (UTMALDG / UBLKCP), mbarrier
operations (SYNCS), fp64 arithmetic (DFMA / DADD / DMUL), the programmatic-dependent-launch pair (ACQBULK = griddepcontrol
.wait, PREEXIT = griddepcontrol.launch_dependents), local-memory traffic (LDL / STL: spills) and global / shared traffic.

usage: python scripts/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "cmpt_eigenex_b200", "lib", "libcmpt_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.splitlines()
names = iter(demangle)
WATCH = ["UTMALDG", "UBLKCP", "UTMAPF", "SYNCS", "DFMA", "DADD", "DMUL", "LDG", "STG", "LDS", "STS", "ATOMG", "RED", "BAR",
         "MEMBAR", "ACQBULK", "PREEXIT", "CCTL", "NANOSLEEP", "SHFL", "LDL", "STL"]
rows = []
cur, cnt, total = None, None, 0
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if cur:
            rows.append((cur, cnt, total))
        cur, cnt, total = next(names), collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        total += 1
        op = m.group(1)
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                cnt[w] += 1
if cur:
    rows.append((cur, cnt, total))
print("# SASS opcode counts per kernel (sm_100a), from `cuobjdump -sass cmpt_eigenex_b200/lib/libcmpt_b200.so`")
print("# UTMALDG = cp.async.bulk.tensor (TMA tile load), UBLKCP = cp.async.bulk (1-D bulk copy), SYNCS = mbarrier ops,")
print("# DFMA/DADD/DMUL = fp64 arithmetic, LDL/STL = local memory (spills), ACQBULK/PREEXIT = griddepcontrol wait / launch_dependents")
print("%-110s %6s  %s" % ("kernel", "instr", "watched opcodes"))
for name, cnt, total in sorted(rows, key=lambda r: r[0]):
    short = re.sub(r"\(.*", "", name)
    print("%-110s %6d  %s" % (short[:110], total, " ".join("%s=%d" % (k, cnt[k]) for k in WATCH if cnt[k])))
