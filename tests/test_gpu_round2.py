"""Round-2 GPU tests, all through the C ABI: the asynchronous lane API and pbh_host_alloc, launch-local tile counters under
concurrent streams and CUDA graphs, the sharded path under REAL NCCL (spawned ranks, needs >= 2 visible GPUs), and the
2^20-item batch bench.py times replayed through the oracle byte for byte."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_batch_equals_oracle_full(gpu_ctx, oracle):
    """Every one of the 2^20 items of ring slot 0 of bench.py (seed 0xB200, D_fullpath, first index 0): proof, status and
    result bytes against the oracle (src/plonk.rs:191-650 restated), for both group algorithms."""
    import torch
    n = 1 << 20
    t, a = gpu_ctx["table"], gpu_ctx["arith"]
    w, r, c, u = t.generate_inputs(n, first_index=0, seed=0xB200, dist=1)
    pt, st = t.prove_batch(w, r, c)
    vt, gt_t = t.verify_batch(pt, c, u, want_gt=True)
    pa, sa = a.prove_batch(w, r, c)
    va, gt_a = a.verify_batch(pa, c, u, want_gt=True)
    t.sync(); a.sync()
    threads = max(1, oracle.hardware_threads())
    hw, hr, hc, hu = (x.cpu().numpy() for x in (w, r, c, u))
    po, so = oracle.prove_batch(hw, hr, hc, threads=threads)
    vo, go = oracle.verify_batch(po, hc, hu, threads=threads)
    for p_, s_, v_, g_ in ((pt, st, vt, gt_t), (pa, sa, va, gt_a)):
        assert np.array_equal(p_.cpu().numpy(), po) and np.array_equal(s_.cpu().numpy(), so)
        assert np.array_equal(v_.cpu().numpy(), vo) and np.array_equal(g_.cpu().numpy(), go)
    # the device generator against the oracle's on a prefix (the oracle needs ~20 prove+verify attempts per item)
    gw, gr, gc, gu = oracle.generate_inputs(8192, first_index=0, seed=0xB200, dist=1, threads=threads)[:4]
    assert np.array_equal(gw, hw[:, :8192]) and np.array_equal(gr, hr[:, :8192]) and np.array_equal(gc, hc[:, :8192]) and np.array_equal(gu, hu[:8192])


def test_lanes_and_host_alloc(gpu_ctx, oracle):
    """pbh_prove_batch_async / pbh_verify_batch_async on two lanes over pbh_host_alloc memory: same bytes as the oracle;
    pageable buffers degrade to the synchronous path with the same bytes; bad lanes and foreign pointers are refused."""
    import pbh_b200
    for algo in ("table", "arith"):
        ctx = gpu_ctx[algo]
        n = 70000 + 48          # ragged: not a multiple of the 256-item tile, a multiple of 16 (TMA path)
        sets = []
        for lane in range(2):
            w, r, c, u, _ = oracle.generate_inputs(n, first_index=lane * n, seed=11, dist=lane)
            bufs = dict(w=ctx.host_alloc((12, n)), r=ctx.host_alloc((9, n)), c=ctx.host_alloc((5, n)), u=ctx.host_alloc((n,)),
                        proof=ctx.host_alloc((27, n)), status=ctx.host_alloc((n,)), result=ctx.host_alloc((n,)), gt=ctx.host_alloc((4, n)))
            bufs["w"][...] = w; bufs["r"][...] = r; bufs["c"][...] = c; bufs["u"][...] = u
            bufs["proof"][...] = 0xEE; bufs["status"][...] = 0xEE; bufs["result"][...] = 0xEE
            sets.append((bufs, (w, r, c, u)))
        expect = []
        for b, (w, r, c, u) in sets:
            po, so = oracle.prove_batch(w, r, c, threads=8)
            expect.append((po, so) + tuple(oracle.verify_batch(po, c, u, threads=8)))
        # PBH_OPT_LANE_MODE: 3 = copy engine both ways (default), 1 = copy-engine upload + in-place stores, 0 = in place
        for mode in (3, 1, 0):
            ctx.set_option(pbh_b200.OPT_LANE_MODE, mode)
            for b, _ in sets:
                b["proof"][...] = 0xEE; b["status"][...] = 0xEE; b["result"][...] = 0xEE; b["gt"][...] = 0xEE
            for rep in range(3):
                for lane, (b, _) in enumerate(sets):
                    ctx.lane_sync(lane)
                    ctx.prove_batch_async(lane, b["w"], b["r"], b["c"], b["proof"], b["status"])
                    ctx.verify_batch_async(lane, b["proof"], b["c"], b["u"], b["result"], gt=b["gt"])
            ctx.sync()
            for (b, _), (po, so, vo, go) in zip(sets, expect):
                assert np.array_equal(b["proof"], po) and np.array_equal(b["status"], so), mode
                assert np.array_equal(b["result"], vo) and np.array_equal(b["gt"], go), mode
        ctx.set_option(pbh_b200.OPT_LANE_MODE, 3)
        # pageable arrays through the same entry points: synchronous, same bytes
        (b, (w, r, c, u)) = sets[1]
        proof = np.full((27, n), 0xEE, np.uint8); status = np.full(n, 0xEE, np.uint8); result = np.full(n, 0xEE, np.uint8)
        ctx.prove_batch_async(3, w, r, c, proof, status)
        ctx.verify_batch_async(3, proof, c, u, result)
        assert np.array_equal(proof, b["proof"]) and np.array_equal(status, b["status"]) and np.array_equal(result, b["result"])
        with pytest.raises(pbh_b200.PbhError):          # lane out of range
            ctx.prove_batch_async(7, w, r, c, proof, status)
        with pytest.raises(pbh_b200.PbhError):
            ctx.host_free(np.zeros(4, np.uint8))
        for b, _ in sets:
            for arr in b.values():
                ctx.host_free(arr)
        assert isinstance(ctx.numa_node, int)


def test_output_arrays_must_be_contiguous(gpu_ctx, oracle):
    """A non-contiguous OUTPUT array used to be filled through a temporary copy and stay untouched."""
    import pbh_b200
    ctx = gpu_ctx["table"]
    w, r, c, u, _ = oracle.generate_inputs(64, seed=3, dist=1)
    bad = np.zeros((64, 27), np.uint8).T          # (27, 64) view with item stride 27
    with pytest.raises(pbh_b200.PbhError):
        ctx.prove_batch(w, r, c, proof=bad)
    ro = np.zeros((27, 64), np.uint8); ro.setflags(write=False)
    with pytest.raises(pbh_b200.PbhError):
        ctx.prove_batch(w, r, c, proof=ro)
    p, s = ctx.prove_batch(np.asfortranarray(w), r, c)         # non-contiguous INPUTS are still copied
    po, so = oracle.prove_batch(w, r, c)
    assert np.array_equal(p, po) and np.array_equal(s, so)


def test_tile_counters_are_launch_local(gpu_ctx, oracle):
    """Two CUDA graphs captured from ONE context and replayed concurrently on different streams, next to direct launches:
    every launch owns its tile-scheduler slot, so no tile is skipped (pbh_capi.cu fresh_tile_counter)."""
    import torch
    ctx = gpu_ctx["table"]
    n = 1 << 18
    dev = torch.device("cuda", 0)
    batches = []
    for k in range(3):
        w, r, c, u = ctx.generate_inputs(n, first_index=k * n, seed=21, dist=1)
        batches.append(dict(w=w, r=r, c=c, u=u, proof=torch.zeros((27, n), dtype=torch.uint8, device=dev), status=torch.full((n,), 77, dtype=torch.uint8, device=dev),
                            result=torch.full((n,), 77, dtype=torch.uint8, device=dev)))
    ctx.sync()
    ref = []
    for b in batches:
        p, s = ctx.prove_batch(b["w"], b["r"], b["c"])
        v = ctx.verify_batch(p, b["c"], b["u"])
        ref.append((p.clone(), s.clone(), v.clone()))
    ctx.sync()
    graphs, streams = [], [torch.cuda.Stream(), torch.cuda.Stream()]
    for b, st in zip(batches[:2], streams):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(4):
                ctx.prove_batch(b["w"], b["r"], b["c"], proof=b["proof"], status=b["status"])
                ctx.verify_batch(b["proof"], b["c"], b["u"], result=b["result"])
        graphs.append(g)
    torch.cuda.synchronize()
    for rep in range(5):
        for g, st in zip(graphs, streams):
            with torch.cuda.stream(st):
                g.replay()
        b = batches[2]
        ctx.prove_batch(b["w"], b["r"], b["c"], proof=b["proof"], status=b["status"])      # direct launches beside the replays
        ctx.verify_batch(b["proof"], b["c"], b["u"], result=b["result"])
    torch.cuda.synchronize(); ctx.sync()
    for b, (p, s, v) in zip(batches, ref):
        assert torch.equal(b["proof"], p) and torch.equal(b["status"], s) and torch.equal(b["result"], v)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, n_total, out_dir):
    sys.path[:0] = [os.path.join(ROOT, "plonk-by-fingers_b200", "python")]
    import torch
    import torch.distributed as dist
    import pbh_b200
    from pbh_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ctx = pbh_b200.Context(device=rank)
    bitmap, digs, total = sharding.run_sharded(n_total, rank, world, lambda lo, cnt: sharding.gpu_compute(ctx, lo, cnt))
    torch.save({"bitmap": bitmap.cpu(), "digs": digs.cpu(), "total": total}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    torch.cuda.synchronize()
    ctx.close()
    dist.destroy_process_group()


def test_run_sharded_under_real_nccl(gpu_ctx, tmp_path):
    """sharding.run_sharded with one process per GPU over NCCL: the gathered bitmap and the summed digests of the N-rank run
    equal what ONE GPU computes for the same global index range."""
    import torch
    import torch.multiprocessing as mp
    from pbh_b200 import sharding
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip(f"needs >= 2 visible GPUs for NCCL ranks, this box shows {torch.cuda.device_count()} "
                    "(shard-count independence on one GPU: test_shard_summaries, test_config4_256m_end_to_end_sharding; "
                    "N-rank equality inside the driver's SCALE run: bench.py check.sharded_equals_single)")
    n_total = (1 << 20) + 4000 + 5
    ctx = gpu_ctx["table"]
    sb, sd, st = sharding.run_sharded(n_total, 0, 1, lambda lo, cnt: sharding.gpu_compute(ctx, lo, cnt))
    port = _free_port()
    mp.spawn(_nccl_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        o = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert torch.equal(o["bitmap"], sb.cpu()) and o["total"] == st and o["digs"].numel() == world


def test_coset_ntt4_exhaustive(gpu_ctx, oracle):
    """pbh_coset_ntt4_batch / pbh_coset_intt4_batch on k H for k = 1 (H), 2 (K1) and 3 (K2), the cosets of
    src/plonk.rs:136-139: ALL 17^4 coefficient tuples.  Forward == Poly::eval (src/poly.rs:71-79) at the four coset points
    k 4^i (the oracle's restatement of eval), inverse(forward(c)) == c, any byte value is reduced like F17::from, host and
    device pointers, ragged sizes."""
    import itertools
    import torch
    import pbh_b200
    ctx = gpu_ctx["table"]
    coeffs = np.ascontiguousarray(np.array(list(itertools.product(range(17), repeat=4)), dtype=np.uint8).T)     # (4, 83521)
    n = coeffs.shape[1]
    assert np.array_equal(ctx.coset_ntt4_batch(coeffs, 1), ctx.ntt4_batch(coeffs))
    for k, coset in ((1, [1, 4, 16, 13]), (2, [2, 8, 15, 9]), (3, [3, 12, 14, 5])):
        ev = ctx.coset_ntt4_batch(coeffs, k)
        for i, point in enumerate(coset):
            arr = np.concatenate([coeffs, np.full((1, n), point, np.uint8)], axis=0)
            assert np.array_equal(ev[i], oracle.poly_eval_batch(arr)), (k, i)
        assert np.array_equal(ctx.coset_intt4_batch(ev, k), coeffs), k
        d = torch.from_numpy(coeffs).cuda()
        ev_d = ctx.coset_ntt4_batch(d, k); back = ctx.coset_intt4_batch(ev_d, k)
        ctx.sync()
        assert np.array_equal(ev_d.cpu().numpy(), ev) and np.array_equal(back.cpu().numpy(), coeffs)
        rng = np.random.default_rng(k)
        raw = rng.integers(0, 256, size=(4, 10007), dtype=np.uint8)
        assert np.array_equal(ctx.coset_ntt4_batch(raw, k), ctx.coset_ntt4_batch(raw % 17, k))
        assert np.array_equal(ctx.coset_intt4_batch(raw, k), ctx.coset_intt4_batch(raw % 17, k))
        assert np.array_equal(ctx.coset_intt4_batch(ctx.coset_ntt4_batch(raw, k), k), raw % 17)
    with pytest.raises(pbh_b200.PbhError):
        ctx.coset_ntt4_batch(coeffs, 4)


def test_poly_mul_all_small_shapes_and_bytes(gpu_ctx, oracle):
    """Schoolbook product (src/poly.rs:205-218) through the size-specialised FP32 kernel: every shape up to 8 x 8 (and the
    general kernel beyond), operands of ANY byte value (F17::from reduces them), ragged and unaligned batches."""
    import torch
    ctx = gpu_ctx["table"]
    rng = np.random.default_rng(77)
    for la in range(1, 10):
        for lb in (1, 2, 3, 6, 7, 8, 9):
            n = 4099 if (la + lb) % 2 else 4096
            a = rng.integers(0, 256, size=(la, n), dtype=np.uint8); b = rng.integers(0, 256, size=(lb, n), dtype=np.uint8)
            a[:, :300] = 255; b[:, :300] = 255; a[:, 300:600] = 0
            exp = oracle.poly_mul_batch(a % 17, b % 17)
            assert np.array_equal(ctx.poly_mul_batch(a, b), exp), (la, lb)
            da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
            got = ctx.poly_mul_batch(da, db); ctx.sync()
            assert np.array_equal(got.cpu().numpy(), exp), (la, lb)
            got = ctx.poly_mul_batch(da[:, 1:], db[:, 1:]); ctx.sync()          # unaligned bases: general kernel
            assert np.array_equal(got.cpu().numpy(), exp[:, 1:]), (la, lb)


def test_multi_device_context(gpu_ctx, oracle):
    """pbh_multi_* (include/pbh_b200.h): every visible GPU (one on a single-GPU box) driven from this one process.  Host-pointer
    prove / verify split over the devices equal the oracle; the sharded pass (in-library NCCL all-gather when there is more
    than one device) returns the bitmap and the digest one context computes alone for the same global index range."""
    import torch
    import pbh_b200
    n_dev = min(torch.cuda.device_count(), 8)
    ctx = gpu_ctx["table"]
    for algo in ("table", "arith"):
        with pbh_b200.MultiContext(n_dev, algo=algo) as mc:
            assert mc.device_count == n_dev
            n = 3 * 256 * n_dev + 77
            w, r, c, u, _ = oracle.generate_inputs(n, seed=31, dist=0, threads=8)
            p, s = mc.prove_batch(w, r, c)
            po, so = oracle.prove_batch(w, r, c, threads=8)
            assert np.array_equal(p, po) and np.array_equal(s, so)
            v, g = mc.verify_batch(p, c, u, want_gt=True)
            vo, go = oracle.verify_batch(po, c, u, threads=8)
            assert np.array_equal(v, vo) and np.array_equal(g, go)
            pk = mc.prove_packed(pbh_b200.pack_witness(w, r, c, u))           # the same shards over packed records
            want = oracle.packed_pack_proofs(po, so)
            assert np.array_equal(np.stack([pk["points_lo"], pk["points_hi"], pk["evals_status"]], axis=1), want)
            assert np.array_equal(mc.verify_packed(pk, pbh_b200.pack_chal_u(c, u)), vo)
            for n_total, first in ((300000 + 40, 0), (4096, 123456 * 8), (77, 8)):
                out = mc.prove_verify_sharded(n_total, first_index=first, seed=0xB200)
                gw, gr, gc, gu = ctx.generate_inputs(n_total, first_index=first, seed=0xB200, dist=1)
                gp, gs = ctx.prove_batch(gw, gr, gc)
                gv = ctx.verify_batch(gp, gc, gu)
                bm = ctx.pack_verdicts(gv)
                dg = int(ctx.digest(gp, first_index=first).item()) & (2**64 - 1)
                ctx.sync()
                assert np.array_equal(out["bitmap"], bm.cpu().numpy()), (algo, n_total)
                assert out["total_digest"] == dg and out["accepted"] == int((gv == 1).sum().item())
                assert sum(int(x) for x in out["digests"]) % 2**64 == dg and out["ms"] > 0
    with pytest.raises(pbh_b200.PbhError):
        pbh_b200.MultiContext([torch.cuda.device_count() + 3])


def test_gt_sweeps_exhaustive(gpu_ctx, oracle):
    """GTP::pow(600) on ALL 101^2 elements of F_101^2 and GTP * GTP on a 101^2 x 40 sample, on the GPU (SURVEY.md section 4;
    src/pbh/gt.rs:33-69 and its vectors :88-97)."""
    ctx = gpu_ctx["arith"]
    allgt = np.ascontiguousarray(np.array([(a, b) for a in range(101) for b in range(101)], dtype=np.uint8).T)
    got = ctx.gt_pow600_batch(allgt)
    for i in range(allgt.shape[1]):
        assert tuple(int(x) for x in got[:, i]) == oracle.gt_pow((int(allgt[0, i]), int(allgt[1, i])), 600), allgt[:, i]
    assert tuple(got[:, 68 * 101 + 47]) == (97, 89)                                  # (68+47u)^600 = 97+89u   src/pbh/gt.rs:96
    rng = np.random.default_rng(3)
    pairs = rng.integers(0, 101, size=(4, 4000), dtype=np.uint8)
    pairs[:, 0] = (26, 97, 93, 76)                                                   # (26+97u)(93+76u) = 97+89u  src/pbh/gt.rs:90
    prod = ctx.gt_mul_batch(pairs)
    assert tuple(prod[:, 0]) == (97, 89)
    for i in range(pairs.shape[1]):
        assert tuple(int(x) for x in prod[:, i]) == oracle.gt_mul((int(pairs[0, i]), int(pairs[1, i])), (int(pairs[2, i]), int(pairs[3, i])))


def test_general_division_and_ragged_subtraction(gpu_ctx, oracle):
    """Poly / Poly for arbitrary divisors (src/poly.rs:230-247; the vectors of src/poly.rs:436-449 as q d + r == n) and the
    reference's += / -= on operands of different lengths (src/poly.rs:165-203), whose subtraction pushes the longer right-hand
    tail UN-NEGATED (Q1): against the oracle's restatement item by item, zero and short operands included."""
    ctx = gpu_ctx["table"]
    rng = np.random.default_rng(21)
    strip = lambda v: (lambda l: l[:max(1, max([k + 1 for k, x in enumerate(l) if x] or [1]))])([int(x) for x in v])
    for ln, ld in ((6, 3), (22, 5), (4, 7), (9, 1), (32, 16)):
        n = 1500
        num = rng.integers(0, 17, size=(ln, n), dtype=np.uint8); den = rng.integers(0, 17, size=(ld, n), dtype=np.uint8)
        den[:, :40] = 0                                    # the zero polynomial: the reference panics
        den[ld // 2:, 40:300] = 0; num[ln // 2:, 200:500] = 0; num[:, 500:520] = 0
        q, r, st = ctx.poly_divrem_batch(num, den)
        for i in range(n):
            try:
                qo, ro = oracle.poly_op(17, "div", strip(num[:, i]), strip(den[:, i]))
                assert st[i] == 0 and strip(q[:, i]) == qo and strip(r[:, i]) == ro, (ln, ld, i)
            except ArithmeticError:
                assert st[i] == 1 and not q[:, i].any() and not r[:, i].any(), (ln, ld, i)
    for la, lb in ((4, 7), (7, 4), (6, 6), (1, 22), (22, 1)):
        n = 1200
        a = rng.integers(0, 17, size=(la, n), dtype=np.uint8); b = rng.integers(0, 17, size=(lb, n), dtype=np.uint8)
        a[la // 2:, :400] = 0; a[:, 400:450] = 0; b[lb // 2:, 300:700] = 0
        for subtract in (False, True):
            out = ctx.poly_addsub_ragged_batch(a, b, subtract=subtract)
            for i in range(n):
                exp = oracle.poly_op(17, "sub" if subtract else "add", strip(a[:, i]), strip(b[:, i]))
                assert strip(out[:, i]) == exp, (la, lb, subtract, i)
    # Q1 by hand: [1] -= [2, 3] gives [16, 3], not [16, 14]
    out = ctx.poly_addsub_ragged_batch(np.array([[1]], np.uint8), np.array([[2], [3]], np.uint8), subtract=True)
    assert out[:, 0].tolist() == [16, 3]


def test_peer_window_replicates_the_summaries(gpu_ctx, oracle):
    """pbh_window_* (include/pbh_b200.h "peer windows"): two ranks of ONE process on cuda:0 (pbh_window_attach_ptrs), each with a
    window of world x bytes_per_rank; the bitmap the verifier writes into its own row and the digest the prover's last block
    writes there must appear at the same offset of the peer's window - on the fused TMA path and on the unaligned fallback -
    and equal what pack_verdicts / digest compute from the result and proof bytes.  Pointers outside the own row are not
    replicated."""
    import torch
    import pbh_b200
    world = 2
    n = 256 * 37 + 64                       # ragged last tile
    nb = (n + 7) // 8
    row = (nb + 8 + 15) // 16 * 16
    bpr = 4 * row
    ranks = [pbh_b200.Context(device=0, algo="table") for _ in range(world)]
    helper = pbh_b200.Context(device=0, algo="table")      # a second stream of rank 0 (pbh_window_share)
    try:
        wins = []
        for r, c in enumerate(ranks):
            w, _h = c.window_create(bpr, r, world)
            wins.append(w)
        for c in ranks:
            c.window_attach_ptrs([c2._window_base for c2 in ranks])
        ranks[0].window_share(helper)
        ref = gpu_ctx["table"]
        for r, c in enumerate(ranks):
            first = r * n
            w, rd, ch, u = ref.generate_inputs(n, first_index=first, seed=99, dist=pbh_b200.DIST_FULLPATH)
            ref.sync()
            proof = torch.empty((27, n), dtype=torch.uint8, device="cuda"); status = torch.empty((n,), dtype=torch.uint8, device="cuda")
            result = torch.empty((n,), dtype=torch.uint8, device="cuda")
            for slot, aligned in ((0, True), (1, False)):
                mine = wins[r][r, slot * row:(slot + 1) * row]
                bitmap, digest = mine[:nb], mine[nb:nb + 8].view(torch.int64)
                if aligned:
                    prover = helper if r == 0 else c
                    prover.prove_digest_batch(w, rd, ch, proof, status, digest, first_index=first)
                    prover.sync()
                    c.verify_bitmap_batch(proof, ch, u, result, bitmap)
                else:                       # planes at an odd offset: the plain-load kernels + separate digest / pack kernels + publish
                    wo = torch.empty((12, n + 16), dtype=torch.uint8, device="cuda")[:, 1:n + 1]; wo.copy_(w)
                    po = torch.empty((27, n + 16), dtype=torch.uint8, device="cuda")[:, 3:n + 3]
                    c.prove_digest_batch(wo, rd, ch, po, status, digest, first_index=first)
                    c.verify_bitmap_batch(po, ch, u, result, bitmap)
                    c.sync()
                    assert torch.equal(po, proof)
                c.sync()
                want_bm = ref.pack_verdicts(result); want_dg = ref.digest(proof, first_index=first)
                ref.sync()
                for holder in range(world):          # the own window and the peer's
                    got = wins[holder][r, slot * row:(slot + 1) * row]
                    assert torch.equal(got[:nb], want_bm), (r, slot, holder)
                    assert int(got[nb:nb + 8].view(torch.int64).item()) == int(want_dg.item()), (r, slot, holder)
            # a summary outside the own row (here: the peer's row of the own window) is written locally only
            other = wins[r][1 - r, 2 * row:3 * row]
            before = wins[1 - r][1 - r, 2 * row:3 * row].clone()
            c.verify_bitmap_batch(proof, ch, u, result, other[:nb])
            c.sync()
            assert torch.equal(other[:nb], want_bm) and torch.equal(wins[1 - r][1 - r, 2 * row:3 * row], before)
        with pytest.raises(pbh_b200.PbhError):
            ranks[0].window_create(24, 0, 2)            # not a multiple of 16
    finally:
        for c in ranks + [helper]:
            c.close()


def test_pageable_staging_by_the_library_and_by_the_driver(gpu_ctx, oracle):
    """PBH_OPT_HOST_STAGE: pageable caller memory staged through the library's page-locked mirrors + copy threads (1, default)
    or by the driver (0) gives the same bytes, for byte planes with odd pitches, records and packed records, over several
    staged chunks and slot reuse (chunk = 2^12 items -> more chunks than staging slots)."""
    import pbh_b200
    ctx = gpu_ctx["table"]
    n = 9 * 4096 + 333
    w, r, c, u, _ = oracle.generate_inputs(n, seed=123, dist=0, threads=8)
    po, so = oracle.prove_batch(w, r, c, threads=8)
    vo, go = oracle.verify_batch(po, c, u, threads=8)
    wide = np.zeros((12, n + 37), np.uint8); wide[:, 5:5 + n] = w          # odd pitch and offset: pageable AND unaligned
    ctx.set_option(pbh_b200.OPT_CHUNK_LOG2, 12)
    try:
        for stage in (1, 0, 1):
            ctx.set_option(pbh_b200.OPT_HOST_STAGE, stage)
            proof = np.full((27, n + 11), 0xEE, np.uint8)
            status = np.full(n, 0xEE, np.uint8)
            ctx.prove_batch(wide[:, 5:5 + n], r, c, proof=proof[:, 3:3 + n], status=status)
            assert np.array_equal(proof[:, 3:3 + n], po) and np.array_equal(status, so), stage
            assert (proof[:, :3] == 0xEE).all() and (proof[:, 3 + n:] == 0xEE).all()
            res, gt = ctx.verify_batch(proof[:, 3:3 + n], c, u, want_gt=True)
            assert np.array_equal(res, vo) and np.array_equal(gt, go), stage
            p2, s2, r2 = ctx.prove_verify_batch(w, r, c, u)
            assert np.array_equal(p2, po) and np.array_equal(s2, so) and np.array_equal(r2, vo), stage
            pk = ctx.prove_packed(pbh_b200.pack_witness(w, r, c, u))
            want = oracle.packed_pack_proofs(po, so)
            assert np.array_equal(np.stack([pk["points_lo"], pk["points_hi"], pk["evals_status"]], axis=1), want), stage
            assert np.array_equal(ctx.verify_packed(pk, pbh_b200.pack_chal_u(c, u)), vo), stage
            recs = ctx.prove_records(pbh_b200.witness_records(w, r, c, u))
            assert np.array_equal(pbh_b200.proof_planes(recs)[0], po), stage
    finally:
        ctx.set_option(pbh_b200.OPT_HOST_STAGE, 1)
        ctx.set_option(pbh_b200.OPT_CHUNK_LOG2, 18)
