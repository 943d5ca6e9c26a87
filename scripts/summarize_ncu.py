#!/usr/bin/env python
"""Turn gpurun_out/<tag>_launches.csv and <tag>_<kernel>.ncu-rep into small text summaries under profiles/.
usage: python scripts/summarize_ncu.py <tag> <kernel-regex> [note]"""
import collections
import csv
import io
import os
import subprocess
import sys

tag, kernel = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, dst = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
os.makedirs(dst, exist_ok=True)

have_list = os.path.exists(os.path.join(src, f"{tag}_launches.csv"))   # a capture of a side kernel may come without a launch list
lines = [l for l in open(os.path.join(src, f"{tag}_launches.csv")) if l.startswith('"')] if have_list else []
rows = list(csv.DictReader(lines)) if have_list else []
agg = collections.OrderedDict()
for r in rows:
    agg.setdefault(r["Kernel Name"].split("(")[0], []).append(float(r["Metric Value"]))
tot = sum(sum(v) for v in agg.values()) or 1.0
# the kernels of a bench step: the default (PBH_ALGO_TABLE = 1) instantiations of the prover and the verifier
step = {k: v for k, v in agg.items() if any(s in k for s in ("prove_f32_tma_kernel<1", "verify_tma_kernel<1", "prove_kernel<1", "verify_kernel<1",
                                                               "pack_verdicts", "digest_kernel"))}
step_tot = sum(sum(v) for v in step.values()) or 1.0
with open(os.path.join(dst, f"{tag}_launches.txt") if have_list else os.devnull, "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, command: bench.py --steps 3 --warmup 3 --no-cpu --ring 2 --e2e-steps 0\n")
    f.write(f"# {note}\n# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
    f.write(f"{'kernel':60s} {'launches':>8s} {'avg_us':>10s} {'share_all_%':>11s} {'share_of_step_%':>15s}\n")
    for k, v in agg.items():
        s2 = f"{100 * sum(v) / step_tot:15.1f}" if k in step else " " * 15
        f.write(f"{k[:60]:60s} {len(v):8d} {sum(v) / len(v) / 1e3:10.1f} {100 * sum(v) / tot:11.1f} {s2}\n")

rep = os.path.join(src, f"{tag}_{kernel}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
data = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = data[0], data[1], data[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size",
        "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg"]
want += [h for h in hdr if h.startswith("smsp__average_warp") or ("warp_issue_stalled" in h and h.endswith("_per_warp_active.pct"))]
with open(os.path.join(dst, f"{tag}_{kernel}_full.txt"), "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on -k regex:{kernel} (from {os.path.basename(rep)})\n# {note}\n")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            f.write(f"{w:80s} {units[i]:16s} {' | '.join(v[i] for v in vals)}\n")
# per-launch DRAM traffic of the profiled kernel, for bench.py's roofline.traffic
import json
def _bytes(name):
    i = hdr.index(name); u = units[i]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    return sum(float(v[i]) for v in vals) / len(vals) * scale
tj = os.path.join(dst, "roofline_traffic.json")
traffic = json.load(open(tj)) if os.path.exists(tj) else {}
def _avg(name):
    i = hdr.index(name)
    return sum(float(v[i]) for v in vals) / len(vals)
if have_list:
  traffic[kernel] = {"tag": tag, "dram_read_bytes": _bytes("dram__bytes_read.sum"), "dram_write_bytes": _bytes("dram__bytes_write.sum"),
                     "issue_active_pct": _avg("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                     "warp_instructions": _avg("smsp__inst_executed.sum"),
                     "note": "ncu --set full, per launch, bench.py --steps 3 --ring 2 (2^20 items); writes of a 28 MB result mostly stay in the 126 MB L2 at capture time"}
  json.dump(traffic, open(tj, "w"), indent=1, sort_keys=True)
  print(open(os.path.join(dst, f"{tag}_launches.txt")).read())
print(open(os.path.join(dst, f"{tag}_{kernel}_full.txt")).read())
