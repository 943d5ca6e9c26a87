"""ncu target, end of round 2: the step's kernels as bench.py launches them (prove + fused digest, verify + fused bitmap), then the
packed-format conversion kernels, each launched twice on 2^20 items (read the second launch of each)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "plonk-by-fingers_b200", "python"))
import numpy as np, torch, pbh_b200
n = 1 << 20
ctx = pbh_b200.Context()
w, r, c, u = ctx.generate_inputs(n, seed=0xB200, dist=1)
dev = w.device
proof = torch.empty((27, n), dtype=torch.uint8, device=dev); status = torch.empty((n,), dtype=torch.uint8, device=dev)
result = torch.empty((n,), dtype=torch.uint8, device=dev)
summary = torch.zeros(n // 8 + 8, dtype=torch.uint8, device=dev)
ctx.sync()
for _ in range(2):
    ctx.prove_digest_batch(w, r, c, proof, status, summary[n // 8:].view(torch.int64), first_index=0)
for _ in range(2):
    ctx.verify_bitmap_batch(proof, c, u, result, summary[:n // 8])
ctx.sync()
pin = torch.from_numpy(pbh_b200.pack_witness(w.cpu().numpy(), r.cpu().numpy(), c.cpu().numpy(), u.cpu().numpy()).view(np.uint8)).to(dev)
cu = torch.from_numpy(pbh_b200.pack_chal_u(c.cpu().numpy(), u.cpu().numpy()).view(np.uint8)).to(dev)
for _ in range(2):
    ctx.unpack_witness_dev(pin, n)
for _ in range(2):
    packed = ctx.pack_proof_dev(proof, status)
for _ in range(2):
    ctx.unpack_proof_dev(packed, cu, n)
ctx.sync(); torch.cuda.synchronize()
print("profile target done")
