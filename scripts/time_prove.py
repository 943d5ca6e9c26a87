"""Times the prove / verify kernels alone (CUDA events on the context's stream) for each launch shape / option."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "plonk-by-fingers_b200", "python"))
import torch, pbh_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
ctx = pbh_b200.Context()
st = ctx.torch_stream()
ring = 8
ins = [ctx.generate_inputs(n, first_index=r * n, seed=0xB200, dist=1) for r in range(ring)]
proof = [torch.empty((27, n), dtype=torch.uint8, device="cuda") for _ in range(ring)]
status = torch.empty((n,), dtype=torch.uint8, device="cuda")
res = torch.empty((n,), dtype=torch.uint8, device="cuda")
ctx.sync()
def timeit(fn, reps=40):
    with torch.cuda.stream(st):
        for k in range(5): fn(k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for k in range(reps): fn(k)
        e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
def prove(k):
    w, r, c, u = ins[k % ring]
    ctx.prove_batch(w, r, c, proof=proof[k % ring], status=status)
def verify(k):
    w, r, c, u = ins[k % ring]
    ctx.verify_batch(proof[k % ring], c, u, result=res)
for spec in (1, 0):
    ctx.set_option(pbh_b200.OPT_SPECIALISE, spec)
    us = timeit(prove)
    print(f"fp32 prover, compile-time circuit constants={spec}: {us:8.1f} us  {n/us/1e3:7.2f} G proofs/s")
ctx.set_option(pbh_b200.OPT_SPECIALISE, 1)
for tma_on in (1, 0):
    ctx.set_option(pbh_b200.OPT_TMA, tma_on)
    us = timeit(prove)
    print(f"fp32 prover, TMA tiles={tma_on}:   {us:8.1f} us  {n/us/1e3:7.2f} G proofs/s  ok={int((status==0).sum())==n}")
for shape in range(0):
    ctx.set_option(pbh_b200.OPT_PROVER_FP32, 1); ctx.set_option(2, shape)
    us = timeit(prove)
    print(f"fp32 prover launch shape {shape}: {us:8.1f} us  {n/us/1e3:7.2f} G proofs/s  ok={int((status==0).sum())==n}")
ctx.set_option(pbh_b200.OPT_PROVER_FP32, 0)
us = timeit(prove); print(f"int32 prover:               {us:8.1f} us  {n/us/1e3:7.2f} G proofs/s")
for tma_on in (1, 0):
    ctx.set_option(pbh_b200.OPT_TMA, tma_on)
    us = timeit(verify); print(f"verify (table), TMA={tma_on}:     {us:8.1f} us  {n/us/1e3:7.2f} G verifies/s")
ctx.set_option(pbh_b200.OPT_TMA, 1)
ctx.set_algo("arith")
us = timeit(verify); print(f"verify (arith), TMA=1:     {us:8.1f} us  {n/us/1e3:7.2f} G verifies/s")
us = timeit(prove); print(f"prove (arith):             {us:8.1f} us  {n/us/1e3:7.2f} G proofs/s")
