#!/usr/bin/env python
"""Fiat-Shamir kernels (pbh_prove_fs_batch_dev / pbh_verify_fs_batch_dev): time per 2^20 items, CUDA events on the
context's stream, D_uniform witnesses (most proofs end in one of the reference's panics) and the
sub-batch of witnesses whose Fiat-Shamir proof exists."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "plonk-by-fingers_b200", "python"))
import torch
import pbh_b200
if os.environ.get('PBH_LIB'):
    pbh_b200.LIB_PATH = os.environ['PBH_LIB']   # experiment: a library built with other launch bounds

n = 1 << 20
ctx = pbh_b200.Context(device=0)
stream = ctx.torch_stream()


def timed(fn, reps=20):
    with torch.cuda.stream(stream):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


w, r, c, u = ctx.generate_inputs(n, seed=1, dist=pbh_b200.DIST_UNIFORM)
proof = torch.empty((27, n), dtype=torch.uint8, device="cuda"); status = torch.empty((n,), dtype=torch.uint8, device="cuda")
chal = torch.empty((6, n), dtype=torch.uint8, device="cuda"); result = torch.empty((n,), dtype=torch.uint8, device="cuda")
for name, algo in (("table", "table"), ("arith", "arith")):
    ctx.set_algo(algo)
    t = timed(lambda: ctx.prove_fs_batch(w, r, proof=proof, status=status, chal=chal))
    ctx.sync()
    ok = status == 0
    print(f"[{name}] prove_fs  D_uniform witnesses: {t*1e3:8.1f} us per 2^20  ({n/t/1e6:7.2f} G/s)  proofs produced: {int(ok.sum())} ({100*float(ok.float().mean()):.1f} %)")
    t = timed(lambda: ctx.verify_fs_batch(proof, result=result, chal=chal))
    ctx.sync()
    print(f"[{name}] verify_fs of that batch (failed items are zero proofs): {t*1e3:8.1f} us per 2^20  ({n/t/1e6:7.2f} G/s)  accepted: {int((result == 1).sum())}")
    # a batch in which every item completes the transcript: tile the successful witnesses
    idx = torch.nonzero(ok).flatten()
    idx = idx.repeat((n + idx.numel() - 1) // idx.numel())[:n]
    w2, r2 = w[:, idx].contiguous(), r[:, idx].contiguous()
    t = timed(lambda: ctx.prove_fs_batch(w2, r2, proof=proof, status=status, chal=chal))
    ctx.sync()
    assert bool((status == 0).all())
    print(f"[{name}] prove_fs  all items complete: {t*1e3:8.1f} us per 2^20  ({n/t/1e6:7.2f} G proofs/s)")
    t = timed(lambda: ctx.prove_fs_batch(w2, r2, proof=proof, status=status, want_chal=False))
    print(f"[{name}] prove_fs  all items complete, challenges not returned (4 compressions): {t*1e3:8.1f} us per 2^20  ({n/t/1e6:7.2f} G proofs/s)")
    t = timed(lambda: ctx.verify_fs_batch(proof, result=result, chal=chal))
    ctx.sync()
    print(f"[{name}] verify_fs all items complete: {t*1e3:8.1f} us per 2^20  ({n/t/1e6:7.2f} G verifies/s)  accepted: {int((result == 1).sum())}")
