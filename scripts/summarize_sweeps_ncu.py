#!/usr/bin/env python
"""Per-kernel summary of an ncu --set full report holding several sweep kernels (last captured launch of each).
usage: python scripts/summarize_sweeps_ncu.py <report.ncu-rep> <items per launch> [note] > profiles/<name>.txt"""
import csv, io, subprocess, sys
rep, items = sys.argv[1], int(sys.argv[2])
note = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
def g(v, name, scale=1.0):
    return float(v[col[name]]) * scale if name in col and v[col[name]] not in ("", "n/a") else float("nan")
def byt(v, name):
    return g(v, name, {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[col[name]], 1))
last = {}
for v in vals:
    last[v[col["Kernel Name"]].split("(")[0]] = v
print(f"# ncu --set full --clock-control none --import-source on, {items} items per launch, last captured launch of each kernel ({rep.split('/')[-1]})")
if note:
    print(f"# {note}")
print("# instr/item = thread instructions per item (smsp__inst_executed.sum x 32 / items); issue / fma / alu / lsu = % of peak sustained active;")
print("# dram = (read + write) bytes per item; times under ncu are cold-cache, compare ratios not absolutes")
print(f"{'kernel':44s} {'us':>8s} {'regs':>5s} {'instr/item':>10s} {'issue%':>7s} {'fma%':>6s} {'alu%':>6s} {'lsu%':>6s} {'warps%':>7s} {'dram B/item':>11s} {'dram%':>6s}")
for k, v in last.items():
    ins = g(v, "smsp__inst_executed.sum") * 32 / items
    print(f"{k[:44]:44s} {g(v, 'gpu__time_duration.sum'):8.1f} {int(g(v, 'launch__registers_per_thread')):5d} {ins:10.1f} "
          f"{g(v, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):7.1f} {g(v, 'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active'):6.1f} "
          f"{g(v, 'sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active'):6.1f} {g(v, 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{g(v, 'sm__warps_active.avg.pct_of_peak_sustained_active'):7.1f} {(byt(v, 'dram__bytes_read.sum') + byt(v, 'dram__bytes_write.sum')) / items:11.2f} "
          f"{g(v, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f}")
