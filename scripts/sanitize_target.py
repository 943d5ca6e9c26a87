"""Small end-to-end run for compute-sanitizer: every kernel family once, TMA and fallback paths, ragged sizes."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "plonk-by-fingers_b200", "python"))
import numpy as np, torch, pbh_b200
for algo in ("table", "arith"):
    ctx = pbh_b200.Context(algo=algo)
    for n in (1, 300, 4096 + 48, 5003):
        pitch = (n + 15) // 16 * 16
        dev = torch.device("cuda", 0)
        w, r, c, u = ctx.generate_inputs(n, seed=n, dist=0)
        mk = lambda planes, src: (lambda t: (t[:, :n].copy_(src), t[:, :n])[1])(torch.zeros((planes, pitch), dtype=torch.uint8, device=dev))
        for (W, R, Cc) in ((w, r, c), (mk(12, w), mk(9, r), mk(5, c))):      # fallback (pitch = n) and TMA (aligned pitch)
            proof = torch.zeros((27, pitch), dtype=torch.uint8, device=dev)[:, :n]
            status = torch.zeros((n,), dtype=torch.uint8, device=dev); result = torch.zeros((n,), dtype=torch.uint8, device=dev)
            digest = torch.zeros((1,), dtype=torch.int64, device=dev); bitmap = torch.zeros(((n + 7) // 8 + 3) // 4 * 4, dtype=torch.uint8, device=dev)
            ctx.prove_digest_batch(W, R, Cc, proof, status, digest, first_index=7)
            ctx.verify_bitmap_batch(proof, Cc, u, result, bitmap[: (n + 7) // 8])
            ctx.verify_batch(proof, Cc, u, want_gt=True)
        hw, hr, hc, hu = (t.cpu().numpy() for t in (w, r, c, u))
        ctx.prove_verify_batch(hw, hr, hc, hu)
    a4 = torch.randint(0, 17, (4, 1003), dtype=torch.uint8, device="cuda")
    ctx.ntt4_batch(a4); ctx.intt4_batch(a4)
    ctx.kzg_commit_batch(torch.randint(0, 17, (7, 1003), dtype=torch.uint8, device="cuda"))
    pts = torch.randint(0, 101, (6, 1003), dtype=torch.uint8, device="cuda"); pts[2] = 0; pts[5] = 0
    ctx.g1_add_batch(pts); ctx.g1_smul_batch(pts[:4]); ctx.pairing_batch(pts[:5])
    ctx.poly_div_zh_batch(torch.randint(0, 17, (22, 1003), dtype=torch.uint8, device="cuda"))
    ctx.poly_mul_batch(torch.randint(0, 17, (6, 1003), dtype=torch.uint8, device="cuda"), torch.randint(0, 17, (7, 1003), dtype=torch.uint8, device="cuda"))
    ctx.sync(); ctx.close()
torch.cuda.synchronize()
print("sanitize target done")
