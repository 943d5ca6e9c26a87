#!/usr/bin/env python
"""Aggregate host <-> device copy rates when 1, 2, 4, 8 GPUs of the box copy AT THE SAME TIME (pinned host memory, one process,
one pair of streams per device): the host-side ceiling that bounds the end-to-end rate of bench.py at N > 1."""
import json
import sys
import time
import torch

nbytes = 128 << 20
ndev = torch.cuda.device_count()
bufs = []
for d in range(ndev):
    dev = torch.device("cuda", d)
    bufs.append(dict(h_in=torch.empty(nbytes, dtype=torch.uint8).pin_memory(), h_out=torch.empty(nbytes, dtype=torch.uint8).pin_memory(),
                     d_in=torch.empty(nbytes, dtype=torch.uint8, device=dev), d_out=torch.ones(nbytes, dtype=torch.uint8, device=dev),
                     s1=torch.cuda.Stream(device=dev), s2=torch.cuda.Stream(device=dev)))


def run(k, h2d, d2h, reps=8):
    for d in range(k):
        torch.cuda.synchronize(d)
    t0 = time.perf_counter()
    for _ in range(reps):
        for d in range(k):
            b = bufs[d]
            if h2d:
                with torch.cuda.stream(b["s1"]):
                    b["d_in"].copy_(b["h_in"], non_blocking=True)
            if d2h:
                with torch.cuda.stream(b["s2"]):
                    b["h_out"].copy_(b["d_out"], non_blocking=True)
    for d in range(k):
        torch.cuda.synchronize(d)
    return (time.perf_counter() - t0) / reps


out = {}
k = 1
while k <= ndev:
    row = {}
    for name, a, b in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
        run(k, a, b, 2)
        t = run(k, a, b)
        row[name] = {"GBps_per_gpu_per_direction": nbytes / t / 1e9, "GBps_aggregate": (a + b) * k * nbytes / t / 1e9}
    out[str(k)] = row
    print(k, "GPUs:", {n: (round(v["GBps_per_gpu_per_direction"], 1), round(v["GBps_aggregate"], 1)) for n, v in row.items()}, file=sys.stderr)
    k *= 2
print(json.dumps(out))
