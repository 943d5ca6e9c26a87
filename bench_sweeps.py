#!/usr/bin/env python
"""Kernel sweep (BASELINE.json configs[3]): NTT/iNTT, Z_H division, KZG commits, G1 and pairing kernels over large
batches resident in HBM.  Prints one JSON line per kernel: items/s, algorithmic GB/s and the fraction of the measured
HBM copy bandwidth (MEASURED_PEAKS.json).  Usage: python bench_sweeps.py [--log2n 24] [--reps 20]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "plonk-by-fingers_b200", "python"))
import torch
import pbh_b200

ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=24)
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
n = 1 << args.log2n
peak = 6558.1
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
ctx = pbh_b200.Context()
st = ctx.torch_stream()
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
rnd = lambda planes, hi: torch.randint(0, hi, (planes, n), dtype=torch.uint8, device=dev, generator=g)

def timeit(fn):
    with torch.cuda.stream(st):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(args.reps): fn()
        e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.reps * 1e-3

def report(name, bytes_per_item, sec, note=""):
    gbs = bytes_per_item * n / sec / 1e9
    print(json.dumps({"kernel": name, "items": n, "us": round(sec * 1e6, 1), "items_per_s": n / sec, "bytes_per_item": bytes_per_item,
                      "GBps": round(gbs, 1), "hbm_frac": round(gbs / peak, 4), "note": note}))

a4 = rnd(4, 17)
report("ntt4", 8, timeit(lambda: ctx.ntt4_batch(a4)), "src/fft.rs:66-106")
report("intt4", 8, timeit(lambda: ctx.intt4_batch(a4)), "src/plonk.rs:177-179")
p22 = rnd(22, 17)
report("poly_div_zh", 44, timeit(lambda: ctx.poly_div_zh_batch(p22)), "22 in + 18 + 4 out; src/poly.rs:230-247")
b22 = rnd(22, 17)
report("poly_add_22", 66, timeit(lambda: ctx.poly_add_batch(p22, b22)), "src/poly.rs:165-176")
a6, b6 = rnd(6, 17), rnd(6, 17)
report("poly_mul_6x6", 23, timeit(lambda: ctx.poly_mul_batch(a6, b6)), "src/poly.rs:205-218")
c7 = rnd(7, 17)
for algo in ("table", "arith"):
    ctx.set_algo(algo)
    report(f"kzg_commit[{algo}]", 10, timeit(lambda: ctx.kzg_commit_batch(c7)), "src/plonk.rs:51-58")
ctx.set_algo("table")
pts = rnd(4, 101); pts[2] = 0
report("g1_smul", 7, timeit(lambda: ctx.g1_smul_batch(pts)), "src/pbh/g1.rs:146-168 (7-bit scalars)")
pq = rnd(5, 101); pq[2] = 0
report("pairing", 7, timeit(lambda: ctx.pairing_batch(pq)), "src/pbh/pairing.rs:12-47 (Miller loop + final exponentiation)")
p8 = rnd(8, 17)
report("poly_scale_7", 15, timeit(lambda: ctx.poly_scale_batch(p8)), "7 coefficient planes + scalar in, 7 out; src/poly.rs:220-228")
report("poly_eval_7", 9, timeit(lambda: ctx.poly_eval_batch(p8)), "7 + point in, 1 out; src/poly.rs:71-79")
report("poly_div_linear_7", 15, timeit(lambda: ctx.poly_div_linear_batch(p8)), "7 + c in, 6 + 1 out; src/poly.rs:230-247, src/plonk.rs:437-442")
p11 = rnd(11, 17)
report("poly_div_linear_10", 21, timeit(lambda: ctx.poly_div_linear_batch(p11)), "the w_z quotient shape of src/plonk.rs:437")
