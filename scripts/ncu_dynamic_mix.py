#!/usr/bin/env python
"""Executed instruction mix and warp-stall sampling of one kernel from an ncu report captured with --import-source on
(the SASS view of `--page source`).  usage: python scripts/ncu_dynamic_mix.py <report.ncu-rep> <items per launch> > profiles/<name>.txt"""
import collections
import csv
import io
import subprocess
import sys

rep, items = sys.argv[1], int(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
i_src, i_ex = hdr.index("Source"), hdr.index("Instructions Executed")
stall_cols = [(h, hdr.index(h)) for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ops, stalls, total = collections.Counter(), collections.Counter(), 0
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break                      # first captured launch only
    if len(r) < len(hdr):
        continue
    tok = r[i_src].split()
    op = (tok[1] if tok[0].startswith("@") else tok[0]).rstrip(";")
    n = int(r[i_ex])
    total += n
    ops[op] += n
    for h, c in stall_cols:
        stalls[h] += int(r[c])
print(f"# {rows[0][1][:120]}")
print(f"# from {rep.split('/')[-1]} (ncu --set full --import-source on), first captured launch, {items} items per launch")
print(f"warp instructions executed: {total}  = {total * 32 / items:.1f} thread instructions per item")
print(f"{'opcode':26s} {'per item':>9s} {'share %':>8s}")
for k, v in ops.most_common(30):
    print(f"{k:26s} {v * 32 / items:9.1f} {100 * v / total:8.1f}")
ts = sum(stalls.values())
print("warp stall sampling (all samples), share of samples %:")
for k, v in stalls.most_common(10):
    print(f"  {k:28s} {100 * v / ts:5.1f}")
