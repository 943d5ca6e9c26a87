"""A/B of the fused kernels at 2^20 and 2^24 items across library BUILDS (diagnostic): python scripts/ab_library_builds.py /path/to/libpbh_b200.so.
Used at the end of round 2 to find that a static first tile and a run-time peer-window branch had made the verifier 9 % slower at
2^24 items (both reverted / turned into a template instantiation)."""
import os, sys, json
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # scripts/ -> repository root
sys.path.insert(0, os.path.join(root, "plonk-by-fingers_b200", "python"))
import torch, pbh_b200
pbh_b200.LIB_PATH = sys.argv[1]
ctx = pbh_b200.Context()
st = ctx.torch_stream()
out = {}
for lg in (20, 24):
    n = 1 << lg
    w, r, c, u = ctx.generate_inputs(n, seed=0xB200, dist=1)
    p, s = ctx.prove_batch(w, r, c)
    v = torch.empty((n,), dtype=torch.uint8, device=w.device)
    for algo in ("table", "arith"):
        ctx.set_algo(algo)
        def run(k):
            with torch.cuda.stream(st):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                for _ in range(k): ctx.verify_batch(p, c, u, result=v)
                e1.record(st)
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / k
        run(3)
        out[f"verify_{algo}_2^{lg}_us"] = round(min(run(10) for _ in range(3)) * 1e3, 2)
        def runp(k):
            with torch.cuda.stream(st):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                for _ in range(k): ctx.prove_batch(w, r, c, proof=p, status=s)
                e1.record(st)
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / k
        runp(2)
        out[f"prove_{algo}_2^{lg}_us"] = round(min(runp(5) for _ in range(3)) * 1e3, 2)
    ctx.set_algo("table")
print(os.path.basename(sys.argv[1]), json.dumps(out))
