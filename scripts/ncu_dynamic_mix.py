#!/usr/bin/env python
"""Executed instruction mix and warp-stall sampling of one kernel from an ncu report captured with --import-source on
(the SASS view of `--page source`).  usage: python scripts/ncu_dynamic_mix.py <report.ncu-rep> <items per launch> [captured launch index, default 0] > profiles/<name>.txt"""
import collections
import csv
import io
import subprocess
import sys

rep, items = sys.argv[1], int(sys.argv[2])
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# one section per captured launch, each starting with a "Kernel Name" row followed by its column header
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
rows = rows[starts[which]:(starts[which + 1] if which + 1 < len(starts) else len(rows))]
hdr = rows[1]
i_src, i_ex = hdr.index("Source"), hdr.index("Instructions Executed")
stall_cols = [(h, hdr.index(h)) for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ops, stalls, total = collections.Counter(), collections.Counter(), 0
i_samp = hdr.index("# Samples") if "# Samples" in hdr else None
hot = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    if i_samp is not None:
        hot.append((int(r[i_samp] or 0), r[i_src].strip()))
    tok = r[i_src].split()
    op = (tok[1] if tok[0].startswith("@") else tok[0]).rstrip(";")
    n = int(r[i_ex])
    total += n
    ops[op] += n
    for h, c in stall_cols:
        stalls[h] += int(r[c])
print(f"# {rows[0][1][:120]}")
print(f"# from {rep.split('/')[-1]} (ncu --set full --import-source on), captured launch {which}, {items} items per launch")
print(f"warp instructions executed: {total}  = {total * 32 / items:.1f} thread instructions per item")
print(f"{'opcode':26s} {'per item':>9s} {'share %':>8s}")
for k, v in ops.most_common(30):
    print(f"{k:26s} {v * 32 / items:9.1f} {100 * v / total:8.1f}")
ts = sum(stalls.values())
print("warp stall sampling (all samples), share of samples %:")
for k, v in stalls.most_common(10):
    print(f"  {k:28s} {100 * v / ts:5.1f}")
if hot:
    tot_s = sum(h[0] for h in hot) or 1
    print("hottest SASS instructions by warp-stall samples (share of all samples %):")
    for n_, src_ in sorted(hot, reverse=True)[:12]:
        print(f"  {100 * n_ / tot_s:5.2f}  {src_[:90]}")
