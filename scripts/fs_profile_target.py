#!/usr/bin/env python
"""Target of the ncu capture of the Fiat-Shamir kernels: a few launches of pbh_prove_fs_batch_dev / pbh_verify_fs_batch_dev
on 2^20 witnesses whose transcript completes (every thread does all five compressions)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "plonk-by-fingers_b200", "python"))
import torch
import pbh_b200

n = 1 << 20
ctx = pbh_b200.Context(device=0)
w, r, c, u = ctx.generate_inputs(n, seed=1, dist=pbh_b200.DIST_UNIFORM)
proof = torch.empty((27, n), dtype=torch.uint8, device="cuda"); status = torch.empty((n,), dtype=torch.uint8, device="cuda")
chal = torch.empty((6, n), dtype=torch.uint8, device="cuda"); result = torch.empty((n,), dtype=torch.uint8, device="cuda")
ctx.prove_fs_batch(w, r, proof=proof, status=status, chal=chal)
ctx.sync()
idx = torch.nonzero(status == 0).flatten()
idx = idx.repeat((n + idx.numel() - 1) // idx.numel())[:n]
w2, r2 = w[:, idx].contiguous(), r[:, idx].contiguous()
for _ in range(3):
    ctx.prove_fs_batch(w2, r2, proof=proof, status=status, chal=chal)
    ctx.verify_fs_batch(proof, result=result, chal=chal)
ctx.sync()
print("ok", int((status == 0).sum()), int((result == 1).sum()))
