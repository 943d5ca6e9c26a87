"""Host-pointer (PCIe inside) throughput of prove+verify for several chunk sizes."""
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
for lg in (15, 16, 17, 18, 19, 20):
    ctx.set_option(pbh_b200.OPT_CHUNK_LOG2, lg)
    def two():
        ctx.prove_batch(hw, hr, hc, proof=hp, status=hs); ctx.verify_batch(hp, hc, hu, result=hv)
    def fused():
        ctx.prove_verify_batch(hw, hr, hc, hu, proof=hp, status=hs, result=hv)
    out = []
    for fn in (two, fused):
        for _ in range(3): fn()
        t0 = time.perf_counter()
        for _ in range(20): fn()
        dt = (time.perf_counter() - t0) / 20
        out.append(dt)
    print(f"chunk 2^{lg}: two calls {out[0]*1e3:6.3f} ms ({n/out[0]/1e6:7.1f} M/s, {92.3/out[0]/1e3:5.1f} GB/s PCIe both ways)   fused {out[1]*1e3:6.3f} ms ({n/out[1]/1e6:7.1f} M/s)")
