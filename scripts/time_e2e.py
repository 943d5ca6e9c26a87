"""Host-pointer (PCIe inside) throughput of prove+verify: chunks staged through device buffers (several chunk sizes)
vs. the kernels running in place on the pinned host buffers."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "plonk-by-fingers_b200", "python"))
import torch, numpy as np, pbh_b200
n = 1 << 20
ctx = pbh_b200.Context()
w, r, c, u = ctx.generate_inputs(n, seed=0xB200, dist=1); ctx.sync()
pin = lambda t: (lambda h: (h.copy_(t), h)[1])(torch.empty(t.shape, dtype=torch.uint8).pin_memory())
hw, hr, hc, hu = [pin(t).numpy() for t in (w, r, c, u)]
hp = torch.empty((27, n), dtype=torch.uint8).pin_memory().numpy(); hs = torch.empty((n,), dtype=torch.uint8).pin_memory().numpy()
hv = torch.empty((n,), dtype=torch.uint8).pin_memory().numpy()
ref = None


def timeit(fn, reps=20):
    for _ in range(3): fn()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps


def row(label):
    global ref
    tp = timeit(lambda: ctx.prove_batch(hw, hr, hc, proof=hp, status=hs))
    tv = timeit(lambda: ctx.verify_batch(hp, hc, hu, result=hv))
    tf = timeit(lambda: ctx.prove_verify_batch(hw, hr, hc, hu, proof=hp, status=hs, result=hv))
    sig = (hp.copy(), hs.copy(), hv.copy())
    if ref is None:
        ref = sig
    same = all(np.array_equal(a, b) for a, b in zip(ref, sig))
    print(f"{label}: prove {tp*1e3:6.3f} ms  verify {tv*1e3:6.3f} ms  two calls {(tp+tv)*1e3:6.3f} ms ({n/(tp+tv)/1e6:7.1f} M/s)"
          f"   fused {tf*1e3:6.3f} ms ({n/tf/1e6:7.1f} M/s)  same bytes: {same}", flush=True)


ctx.set_option(pbh_b200.OPT_HOST_DIRECT, 0)
for lg in (17, 18, 19):
    ctx.set_option(pbh_b200.OPT_CHUNK_LOG2, lg)
    row(f"staged  chunk 2^{lg}")
ctx.set_option(pbh_b200.OPT_CHUNK_LOG2, 18)
ctx.set_option(pbh_b200.OPT_HOST_DIRECT, 1)
row("in place on pinned host buffers (fused call stays staged)")
