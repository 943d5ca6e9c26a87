"""ncu target: every sweep kernel of BASELINE.json configs[3] twice on 2^LOG2N items (first call warms, second is the one to read).
usage: python scripts/sweep_profile_target.py [log2n]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "plonk-by-fingers_b200", "python"))
import torch, pbh_b200
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 22)
ctx = pbh_b200.Context()
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
rnd = lambda planes, hi: torch.randint(0, hi, (planes, n), dtype=torch.uint8, device=dev, generator=g)
a4, p22, a6, b6, c7, p8, p11 = rnd(4, 17), rnd(22, 17), rnd(6, 17), rnd(6, 17), rnd(7, 17), rnd(8, 17), rnd(11, 17)
pts = rnd(4, 101); pts[2] = 0
pq = rnd(5, 101); pq[2] = 0
for rep in range(2):
    ctx.ntt4_batch(a4); ctx.intt4_batch(a4); ctx.poly_div_zh_batch(p22); ctx.poly_add_batch(p22, p22)
    ctx.poly_mul_batch(a6, b6)
    ctx.set_algo("arith"); ctx.kzg_commit_batch(c7); ctx.set_algo("table"); ctx.kzg_commit_batch(c7)
    ctx.g1_smul_batch(pts); ctx.pairing_batch(pq)
    ctx.poly_scale_batch(p8); ctx.poly_eval_batch(p8); ctx.poly_div_linear_batch(p8); ctx.poly_div_linear_batch(p11)
    if hasattr(ctx, "coset_ntt4_batch"):
        ctx.coset_ntt4_batch(a4, 2); ctx.coset_intt4_batch(a4, 3)
ctx.sync(); torch.cuda.synchronize()
print("sweep profile target done")
