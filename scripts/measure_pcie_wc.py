#!/usr/bin/env python
"""Does write-combined page-locked memory (pbh_host_alloc_input) raise the host -> device rate when many GPUs copy at once?
H2D from ordinary pinned memory against H2D from write-combined memory, alone and with a D2H stream running beside it, on
1 .. all GPUs of the box at the same time."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "plonk-by-fingers_b200", "python"))
import numpy as np, torch, pbh_b200
try:
    from cuda.bindings import runtime as cudart
except Exception:
    from cuda import cudart
H2D, D2H = cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost

nbytes = 128 << 20
ndev = torch.cuda.device_count()
B = []
for d in range(ndev):
    dev = torch.device("cuda", d)
    ctx = pbh_b200.Context(device=d)
    wc = ctx.host_alloc(nbytes, write_combined=True)      # numpy arrays over the library's page-locked memory
    pl = ctx.host_alloc(nbytes)
    out = ctx.host_alloc(nbytes)
    wc[...] = 1; pl[...] = 1
    B.append(dict(ctx=ctx, wc=wc, pl=pl, out=out, d_in=torch.empty(nbytes, dtype=torch.uint8, device=dev), d_out=torch.ones(nbytes, dtype=torch.uint8, device=dev),
                  s1=torch.cuda.Stream(device=dev), s2=torch.cuda.Stream(device=dev)))


def run(k, src, d2h, reps=8):
    for d in range(k):
        torch.cuda.synchronize(d)
    t0 = time.perf_counter()
    for _ in range(reps):
        for d in range(k):
            b = B[d]
            cudart.cudaSetDevice(d)
            (e,) = cudart.cudaMemcpyAsync(b["d_in"].data_ptr(), b[src].ctypes.data, nbytes, H2D, b["s1"].cuda_stream)
            assert int(e) == 0, e
            if d2h:
                (e,) = cudart.cudaMemcpyAsync(b["out"].ctypes.data, b["d_out"].data_ptr(), nbytes, D2H, b["s2"].cuda_stream)
                assert int(e) == 0, e
    for d in range(k):
        torch.cuda.synchronize(d)
    return (time.perf_counter() - t0) / reps


res = {}
k = 1
while k <= ndev:
    row = {}
    for name, src, d2h in (("h2d_pinned", "pl", False), ("h2d_write_combined", "wc", False), ("h2d_pinned_beside_d2h", "pl", True), ("h2d_write_combined_beside_d2h", "wc", True)):
        run(k, src, d2h, 2)
        t = run(k, src, d2h)
        row[name] = round(nbytes / t / 1e9, 1)
    res[str(k)] = row
    print(k, "GPUs, GB/s per GPU per direction:", row, file=sys.stderr)
    k *= 2
print(json.dumps(res))
