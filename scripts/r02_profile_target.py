"""ncu target for the fused kernels: each launched twice on 2^20 items (the second launch of each is the one to read).
Launch order (pairs): prove TABLE, verify TABLE, prove ARITH, verify ARITH, prove_fs, verify_fs, prove TABLE on D_uniform."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "plonk-by-fingers_b200", "python"))
import torch, pbh_b200
n = 1 << 20
ctx = pbh_b200.Context()
w, r, c, u = ctx.generate_inputs(n, seed=0xB200, dist=1)
wu, ru, cu, uu = ctx.generate_inputs(n, seed=0xB200, dist=0)
dev = w.device
proof = torch.empty((27, n), dtype=torch.uint8, device=dev); status = torch.empty((n,), dtype=torch.uint8, device=dev)
result = torch.empty((n,), dtype=torch.uint8, device=dev); chal6 = torch.empty((6, n), dtype=torch.uint8, device=dev)
ctx.sync()
for algo in ("table", "arith"):
    ctx.set_algo(algo)
    for _ in range(2):
        ctx.prove_batch(w, r, c, proof=proof, status=status)
    for _ in range(2):
        ctx.verify_batch(proof, c, u, result=result)
ctx.set_algo("table")
for _ in range(2):
    ctx.prove_fs_batch(w, r, proof=proof, status=status, chal=chal6)
for _ in range(2):
    ctx.verify_fs_batch(proof, result=result, chal=chal6)
for _ in range(2):
    ctx.prove_batch(wu, ru, cu, proof=proof, status=status)
ctx.sync(); torch.cuda.synchronize()
print("profile target done")
