#!/usr/bin/env python
"""PCIe copy rates of the box (pinned host memory): H2D alone, D2H alone, both directions at once."""
import time
import torch

nbytes = 256 << 20
h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
d_out = torch.ones(nbytes, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


for name, a, b in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
    run(a, b, 2)
    t = run(a, b)
    print(f"{name}: {nbytes / t / 1e9:.1f} GB/s per direction, {(a + b) * nbytes / t / 1e9:.1f} GB/s total")
