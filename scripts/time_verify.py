"""Times the TABLE verifier with its F_17 scalar work on the FP32 pipes (PBH_OPT_VERIFIER_FP32 = 1) and in int32 (the default)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "plonk-by-fingers_b200", "python"))
import torch, pbh_b200
n = 1 << 20
ctx = pbh_b200.Context()
st = ctx.torch_stream()
ring = 8
ins = [ctx.generate_inputs(n, first_index=r * n, seed=0xB200, dist=1) for r in range(ring)]
proof = [torch.empty((27, n), dtype=torch.uint8, device="cuda") for _ in range(ring)]
status = torch.empty((n,), dtype=torch.uint8, device="cuda")
res = torch.empty((n,), dtype=torch.uint8, device="cuda")
for k in range(ring):
    ctx.prove_batch(ins[k][0], ins[k][1], ins[k][2], proof=proof[k], status=status)
ctx.sync()
def timeit(fn, reps=80):
    with torch.cuda.stream(st):
        for k in range(8): fn(k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for k in range(reps): fn(k)
        e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
def verify(k):
    ctx.verify_batch(proof[k % ring], ins[k % ring][2], ins[k % ring][3], result=res)
for rep in range(2):
    for fp32 in (1, 0):
        ctx.set_option(pbh_b200.OPT_VERIFIER_FP32, fp32)
        us = timeit(verify)
        print(f"verify (table), scalar work in {'fp32' if fp32 else 'int32'}: {us:7.2f} us  {n/us/1e3:6.2f} G verifies/s  accepted={int((res==1).sum())}")
