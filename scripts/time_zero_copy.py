#!/usr/bin/env python
"""Experiment: let the tile kernels read their inputs from, and write their outputs to, PINNED HOST memory directly
(UVA: a pinned host pointer is a valid device pointer), instead of staging chunks through device buffers.

Prints time per 2^20-item prove / verify call for (a) the staged host-pointer entry points and (b) the device entry
points handed pinned host pointers, and checks that both produce the same bytes."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "plonk-by-fingers_b200", "python"))
import numpy as np
import torch
import pbh_b200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
reps = 10
ctx = pbh_b200.Context(device=0)
w, r, c, u = ctx.generate_inputs(n, first_index=0, seed=0xB200, dist=pbh_b200.DIST_FULLPATH)
ctx.sync()
torch.cuda.synchronize()


def pinned(t):
    h = torch.empty(t.shape, dtype=torch.uint8).pin_memory()
    h.copy_(t)
    return h


hw, hr, hc, hu = pinned(w), pinned(r), pinned(c), pinned(u)
hp_a = torch.empty((27, n), dtype=torch.uint8).pin_memory()
hs_a = torch.empty((n,), dtype=torch.uint8).pin_memory()
hp_b = torch.zeros((27, n), dtype=torch.uint8).pin_memory()
hs_b = torch.zeros((n,), dtype=torch.uint8).pin_memory()
hres_a = torch.empty((n,), dtype=torch.uint8).pin_memory()
hres_b = torch.zeros((n,), dtype=torch.uint8).pin_memory()
lib, h = ctx.lib, ctx.h


def staged_prove():
    ctx.prove_batch(hw.numpy(), hr.numpy(), hc.numpy(), proof=hp_a.numpy(), status=hs_a.numpy())


def staged_verify():
    ctx.verify_batch(hp_a.numpy(), hc.numpy(), hu.numpy(), result=hres_a.numpy())


def direct_prove():
    rc = lib.pbh_prove_batch_dev(h, C.c_size_t(n), C.c_void_p(hw.data_ptr()), C.c_size_t(n), C.c_void_p(hr.data_ptr()), C.c_size_t(n),
                                 C.c_void_p(hc.data_ptr()), C.c_size_t(n), C.c_void_p(hp_b.data_ptr()), C.c_size_t(n), C.c_void_p(hs_b.data_ptr()))
    assert rc == 0, rc
    ctx.sync()


def direct_verify():
    rc = lib.pbh_verify_batch_dev(h, C.c_size_t(n), C.c_void_p(hp_b.data_ptr()), C.c_size_t(n), C.c_void_p(hc.data_ptr()), C.c_size_t(n),
                                  C.c_void_p(hu.data_ptr()), C.c_void_p(hres_b.data_ptr()), C.c_void_p(None), C.c_size_t(0))
    assert rc == 0, rc
    ctx.sync()


def timeit(fn):
    for _ in range(3):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


res = {}
for name, fn in (("staged_prove", staged_prove), ("staged_verify", staged_verify), ("direct_prove", direct_prove), ("direct_verify", direct_verify)):
    res[name] = timeit(fn)
    print(f"{name:14s} {res[name]:.3f} ms per {n} items", flush=True)
print("same proof bytes:", bool(torch.equal(hp_a, hp_b)), " same status:", bool(torch.equal(hs_a, hs_b)), " same result:", bool(torch.equal(hres_a, hres_b)))
for k in ("staged", "direct"):
    t = res[k + "_prove"] + res[k + "_verify"]
    print(f"{k}: prove+verify {t:.3f} ms -> {n / t / 1e6:.3f} G proof+verify/s")

# which direction is the slow one: inputs on the device + outputs in pinned host memory, and the reverse
dp = torch.empty((27, n), dtype=torch.uint8, device="cuda")
ds = torch.empty((n,), dtype=torch.uint8, device="cuda")
dres = torch.empty((n,), dtype=torch.uint8, device="cuda")


def prove_ptrs(a, b, c_, p, s):
    rc = lib.pbh_prove_batch_dev(h, C.c_size_t(n), C.c_void_p(a.data_ptr()), C.c_size_t(n), C.c_void_p(b.data_ptr()), C.c_size_t(n),
                                 C.c_void_p(c_.data_ptr()), C.c_size_t(n), C.c_void_p(p.data_ptr()), C.c_size_t(n), C.c_void_p(s.data_ptr()))
    assert rc == 0, rc
    ctx.sync()


def verify_ptrs(p, c_, u_, res_):
    rc = lib.pbh_verify_batch_dev(h, C.c_size_t(n), C.c_void_p(p.data_ptr()), C.c_size_t(n), C.c_void_p(c_.data_ptr()), C.c_size_t(n),
                                  C.c_void_p(u_.data_ptr()), C.c_void_p(res_.data_ptr()), C.c_void_p(None), C.c_size_t(0))
    assert rc == 0, rc
    ctx.sync()


torch.cuda.synchronize()
for name, fn, nbytes in (("prove  in=dev  out=host", lambda: prove_ptrs(w, r, c, hp_b, hs_b), 28 * n),
                         ("prove  in=host out=dev ", lambda: prove_ptrs(hw, hr, hc, dp, ds), 26 * n),
                         ("prove  in=dev  out=dev ", lambda: prove_ptrs(w, r, c, dp, ds), 0),
                         ("verify in=host out=dev ", lambda: verify_ptrs(hp_b, hc, hu, dres), 33 * n),
                         ("verify in=dev  out=host", lambda: verify_ptrs(dp, c, u, hres_b), n)):
    t = timeit(fn)
    print(f"{name}: {t:.3f} ms" + (f"  ({nbytes / t / 1e6:.1f} GB/s over PCIe)" if nbytes else ""), flush=True)

# do SM zero-copy traffic and copy-engine DMA in the opposite direction overlap on the link?
side = torch.cuda.Stream()
big_h = torch.empty(26 * n, dtype=torch.uint8).pin_memory()
big_d = torch.empty(26 * n, dtype=torch.uint8, device="cuda")
big_h2 = torch.empty(28 * n, dtype=torch.uint8).pin_memory()
big_d2 = torch.empty(28 * n, dtype=torch.uint8, device="cuda")


def dma_h2d_plus_zero_copy_writes():
    with torch.cuda.stream(side):
        big_d.copy_(big_h, non_blocking=True)
    prove_ptrs(w, r, c, hp_b, hs_b)
    side.synchronize()


def zero_copy_reads_plus_dma_d2h():
    with torch.cuda.stream(side):
        big_h2.copy_(big_d2, non_blocking=True)
    prove_ptrs(hw, hr, hc, dp, ds)
    side.synchronize()


for name, fn in (("DMA H2D 26 B/item  +  kernel in=dev out=host", dma_h2d_plus_zero_copy_writes),
                 ("kernel in=host out=dev  +  DMA D2H 28 B/item", zero_copy_reads_plus_dma_d2h)):
    print(f"{name}: {timeit(fn):.3f} ms", flush=True)
