#!/usr/bin/env python
"""Kernel sweep (BASELINE.json configs[3]): NTT/iNTT (+ the coset variants), Z_H division, KZG commits, polynomial, G1 and
pairing kernels over large batches resident in HBM.  One JSON record per kernel: items/s, algorithmic GB/s, the fraction
of the measured HBM copy bandwidth (MEASURED_PEAKS.json) and, for the kernels that are bound by instruction issue, the
fraction of the issue slots (committed ncu instruction counts x items / this run's CUDA-event time / this run's peak).
bench.py imports run_sweeps() and prints the table in its line (`sweeps`).
Usage: python bench_sweeps.py [--log2n 24] [--reps 20]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "plonk-by-fingers_b200", "python"))


def _instr():
    try:
        with open(os.path.join(ROOT, "profiles", "instr_per_item.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def run_sweeps(ctx, log2n, reps=20, hbm_peak=6558.1, peaks=None):
    import torch
    n = 1 << log2n
    st = ctx.torch_stream()
    dev = torch.device("cuda", ctx.device)
    g = torch.Generator(device=dev); g.manual_seed(1)
    rnd = lambda planes, hi: torch.randint(0, hi, (planes, n), dtype=torch.uint8, device=dev, generator=g)
    instr = _instr()
    out = []

    def timeit(fn):
        with torch.cuda.stream(st):
            for _ in range(3): fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(reps): fn()
            e1.record(st)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e-3

    def report(name, bytes_per_item, fn, note=""):
        sec = timeit(fn)
        gbs = bytes_per_item * n / sec / 1e9
        rec = {"kernel": name, "items": n, "us": round(sec * 1e6, 1), "items_per_s": n / sec, "bytes_per_item": bytes_per_item,
               "GBps": round(gbs, 1), "hbm_frac": round(gbs / hbm_peak, 4), "note": note}
        fr = {"hbm": gbs / hbm_peak}
        ins = instr.get("sweep_" + name)
        if ins and peaks and peaks.get("issue_thread_instr_per_s"):
            rec["instr_per_item_static"] = ins.get("total")
            fr["issue"] = ins["total"] * n / sec / peaks["issue_thread_instr_per_s"]
            rec["issue_frac"] = round(fr["issue"], 4)
        rec["bound"] = max(fr, key=fr.get)
        rec["frac"] = round(fr[rec["bound"]], 4)
        out.append(rec)
        torch.cuda.empty_cache()

    def curve_points(extra_planes, hi):
        """planes x, y, inf = 0 of random points ON the curve y^2 = x^3 + 3 over F_101 (the group kernels take group elements),
        followed by `extra_planes` random planes below `hi`"""
        ys = {}
        for x in range(101):
            for y in range(101):
                if (y * y - x * x * x - 3) % 101 == 0:
                    ys.setdefault(x, []).append(y)
        xs = torch.tensor(sorted(ys), dtype=torch.uint8, device=dev)
        y0 = torch.zeros(101, dtype=torch.uint8, device=dev); y1 = torch.zeros(101, dtype=torch.uint8, device=dev)
        for x, v in ys.items():
            y0[x], y1[x] = v[0], v[-1]
        pick = torch.randint(0, xs.numel(), (n,), device=dev, generator=g)
        x = xs[pick]
        sign = torch.randint(0, 2, (n,), device=dev, generator=g).bool()
        y = torch.where(sign, y1[x.long()], y0[x.long()])
        t = torch.empty((3 + extra_planes, n), dtype=torch.uint8, device=dev)
        t[0], t[1], t[2] = x, y, 0
        for k in range(extra_planes):
            t[3 + k] = torch.randint(0, hi, (n,), dtype=torch.uint8, device=dev, generator=g)
        return t

    a4 = rnd(4, 17)
    report("ntt4", 8, lambda: ctx.ntt4_batch(a4), "src/fft.rs:66-106")
    report("intt4", 8, lambda: ctx.intt4_batch(a4), "src/plonk.rs:177-179")
    if hasattr(ctx, "coset_ntt4_batch"):
        report("coset_ntt4[k1]", 8, lambda: ctx.coset_ntt4_batch(a4, 2), "evaluations on k1 H, k1 = 2 (src/plonk.rs:136-139)")
        report("coset_intt4[k2]", 8, lambda: ctx.coset_intt4_batch(a4, 3), "interpolation from k2 H, k2 = 3")
    del a4
    p22 = rnd(22, 17)
    report("poly_div_zh", 44, lambda: ctx.poly_div_zh_batch(p22), "22 in + 18 + 4 out; src/poly.rs:230-247")
    report("poly_add_22", 66, lambda: ctx.poly_add_batch(p22, p22), "src/poly.rs:165-176")
    del p22
    a6, b6 = rnd(6, 17), rnd(6, 17)
    report("poly_mul_6x6", 23, lambda: ctx.poly_mul_batch(a6, b6), "src/poly.rs:205-218")
    del a6, b6
    c7 = rnd(7, 17)
    for algo in ("table", "arith"):
        ctx.set_algo(algo)
        report(f"kzg_commit[{algo}]", 10, lambda: ctx.kzg_commit_batch(c7), "src/plonk.rs:51-58")
    ctx.set_algo("table")
    del c7
    pts = curve_points(1, 101)
    report("g1_smul", 7, lambda: ctx.g1_smul_batch(pts), "src/pbh/g1.rs:146-168 (7-bit scalars, points on the curve)")
    del pts
    pq = curve_points(2, 101)
    report("pairing", 7, lambda: ctx.pairing_batch(pq), "src/pbh/pairing.rs:12-47 (Miller loop + final exponentiation)")
    del pq
    p8 = rnd(8, 17)
    report("poly_scale_7", 15, lambda: ctx.poly_scale_batch(p8), "7 coefficient planes + scalar in, 7 out; src/poly.rs:220-228")
    report("poly_eval_7", 9, lambda: ctx.poly_eval_batch(p8), "7 + point in, 1 out; src/poly.rs:71-79")
    report("poly_div_linear_7", 15, lambda: ctx.poly_div_linear_batch(p8), "7 + c in, 6 + 1 out; src/poly.rs:230-247, src/plonk.rs:437-442")
    del p8
    p11 = rnd(11, 17)
    report("poly_div_linear_10", 21, lambda: ctx.poly_div_linear_batch(p11), "the w_z quotient shape of src/plonk.rs:437")
    del p11
    if hasattr(ctx, "unpack_witness_dev"):
        # packed wire format conversions (include/pbh_b200.h): valid random records built on the device
        words = torch.stack([torch.randint(0, 17 ** 7, (n,), dtype=torch.int64, device=dev, generator=g) for _ in range(3)] +
                            [torch.randint(0, 17 ** 6, (n,), dtype=torch.int64, device=dev, generator=g)], dim=1).to(torch.int32)
        pin = words.contiguous().view(torch.uint8).reshape(-1)
        del words
        report("unpack_witness", 43, lambda: ctx.unpack_witness_dev(pin, n), "16-byte packed prover input -> 27 byte planes")
        cu = pin.view(torch.int32).reshape(n, 4)[:, 3].contiguous().view(torch.uint8).reshape(-1)
        del pin
        pts9 = torch.randint(0, 102 ** 9, (n,), dtype=torch.int64, device=dev, generator=g)
        ev = torch.randint(0, 17 ** 7, (n,), dtype=torch.int64, device=dev, generator=g)
        rec = torch.stack([pts9 & 0xFFFFFFFF, pts9 >> 32, ev], dim=1).to(torch.int32)
        del pts9, ev
        packed = rec.contiguous().view(torch.uint8).reshape(-1)
        del rec
        report("unpack_proof", 50, lambda: ctx.unpack_proof_dev(packed, cu, n), "12-byte packed proof + 4-byte challenge word -> 27 + 1 + 5 + 1 byte planes")
        planes, status, _c, _u = ctx.unpack_proof_dev(packed, cu, n)
        del packed, cu, _c, _u
        report("pack_proof", 40, lambda: ctx.pack_proof_dev(planes, status), "27 proof planes + status -> 12-byte packed proof")
        del planes, status
    return out


if __name__ == "__main__":
    import torch
    import pbh_b200
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=24)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    peak = 6558.1
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    ctx = pbh_b200.Context()
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    peaks = {"issue_thread_instr_per_s": sm * 128 * 1.965e9}
    for rec in run_sweeps(ctx, args.log2n, args.reps, peak, peaks):
        print(json.dumps(rec))
