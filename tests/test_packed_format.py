"""The packed wire format of include/pbh_b200.h (16-byte prover inputs, 12-byte proofs, 4-byte challenge words).

The reference has no serialisation (src/plonk.rs:61: Proof derives Debug, PartialEq only), so the format itself is pinned by
construction: oracle/oracle.py holds an independent numpy / Python-integer restatement of the header's text, and the
product's codec (host helpers on the CPU, kernels on the GPU) must agree with it word for word.  What IS pinned by the
reference is the content: packed prove / verify must equal Plonk::prove / Plonk::verify (src/plonk.rs:191-650) of the oracle
on the decoded values, byte for byte, for every status class.
"""
import numpy as np
import pytest


def _words(pk):
    return np.stack([pk["points_lo"], pk["points_hi"], pk["evals_status"]], axis=1)


def test_point_codes_are_the_curve(oracle):
    """102 codes: the identity and the 101 finite points of y^2 = x^3 + 3 over F_101, in (x, y) order; every one is on the
    curve for the oracle's G1P::in_curve (src/pbh/g1.rs:63-81) and the generator (1, 2) is code 1."""
    pts = oracle.packed_point_codes()
    assert len(pts) == 102 and pts[0] == (0, 0) and pts[1] == (1, 2) and pts[2] == (1, 99)
    assert pts[1:] == sorted(pts[1:]) and len(set(pts[1:])) == 101
    for x, y in pts[1:]:
        assert oracle.g1_in_curve((x, y))
    # the whole group is reachable: multiples of a generator of the order-102 group cover all codes
    seen = set()
    for x, y in pts[1:]:
        seen.add((x, y))
    assert len(seen) == 101


def test_host_codec_equals_the_restatement(product_lib, oracle):
    import pbh_b200 as P
    rng = np.random.default_rng(2026)
    n = 3000
    w, r, c, u, _ = oracle.generate_inputs(n, seed=5, dist=0, threads=4)
    pk = P.pack_witness(w, r, c, u)
    assert pk.dtype.itemsize == 16 and np.array_equal(pk["w"], oracle.packed_pack_witness(w, r, c, u))
    for got, want in zip(P.unpack_witness(pk), (w, r, c, u)):
        assert np.array_equal(got, want)
    assert np.array_equal(P.pack_chal_u(c, u), oracle.packed_pack_chal_u(c, u))
    assert np.array_equal(P.pack_chal_u(c, u), pk["w"][:, 3])          # word 3 of a packed witness IS its challenge word
    # words outside the format: the top digit saturates (a byte >= 17 comes out), same on both sides
    rw = rng.integers(0, 2**32, (700, 4), dtype=np.uint64).astype("<u4")
    rw[:4] = [[0, 0, 0, 0], [0xFFFFFFFF] * 4, [17**7 - 1] * 3 + [17**6 - 1], [17**7] * 3 + [17**6]]
    for got, want in zip(P.unpack_witness(rw.view(P.PACKED_WITNESS).reshape(-1)), oracle.packed_unpack_witness(rw)):
        assert np.array_equal(got, want)
    cw, cu_ = oracle.packed_unpack_chal_u(rw[:, 3])
    assert np.array_equal(P.unpack_witness(rw.view(P.PACKED_WITNESS).reshape(-1))[2], cw)
    # a value >= 17 has no packed form
    bad = w.copy(); bad[3, 7] = 17
    with pytest.raises(P.PbhError):
        P.pack_witness(bad, r, c, u)

    # proofs of every status class, round trip
    po, so = oracle.prove_batch(w, r, c, threads=4)
    assert len(set(so.tolist())) >= 4
    pp = P.pack_proofs(po, so)
    assert pp.dtype.itemsize == 12 and np.array_equal(_words(pp), oracle.packed_pack_proofs(po, so))
    back, st = P.unpack_proofs(pp)
    assert np.array_equal(back, po) and np.array_equal(st, so)
    ob, os_ = oracle.packed_unpack_proofs(_words(pp))
    assert np.array_equal(ob, po) and np.array_equal(os_, so)
    failed = so != 0
    assert failed.any() and not _words(pp)[failed][:, :2].any() and ((pp["evals_status"][failed] & 0x1FFFFFFF) == 0).all()

    # every point code in every position, identities included
    pts = oracle.packed_point_codes()
    m = 102 * 9
    planes = np.zeros((27, m), np.uint8)
    planes[0:18:2] = 1; planes[1:18:2] = 2
    for k in range(9):
        for code in range(102):
            i = k * 102 + code
            planes[2 * k, i], planes[2 * k + 1, i] = pts[code]
            if code == 0:
                planes[18 + k // 8, i] |= 1 << (k % 8)
    planes[20:27] = rng.integers(0, 17, (7, m), dtype=np.uint8)
    pp = P.pack_proofs(planes)
    assert np.array_equal(_words(pp), oracle.packed_pack_proofs(planes)) and (pp["evals_status"] >> 29 == 0).all()
    back, st = P.unpack_proofs(pp)
    assert np.array_equal(back, planes) and not st.any()

    # proofs the format cannot express get status code 6 (PBH_ST_UNREPRESENTABLE) and zero payload
    odd = np.repeat(planes[:, 5:6], 6, axis=1)
    odd[0, 0], odd[1, 0] = 1, 3              # off the curve
    odd[2, 1] = 101                          # coordinate outside the field
    odd[18, 2] |= 1; odd[0, 2], odd[1, 2] = 1, 2   # flagged identity with coordinates (Q9)
    odd[22, 3] = 17                          # evaluation outside the field
    odd[19, 4] |= 0x02                       # undefined flag bit
    odd[3, 5] = 101                          # y == 101 must not alias 101 - 0
    pp = P.pack_proofs(odd)
    assert np.array_equal(_words(pp), oracle.packed_pack_proofs(odd))
    assert (pp["evals_status"] == 6 << 29).all() and not pp["points_lo"].any() and not pp["points_hi"].any()
    back, st = P.unpack_proofs(pp)
    assert not back.any() and (st == P.ST_UNREPRESENTABLE).all()
    # arbitrary words decode identically (ninth digit >= 102 -> the point (0, 0) without its flag)
    rp = rng.integers(0, 2**32, (900, 3), dtype=np.uint64).astype("<u4")
    rp[:600, 2] &= 0x1FFFFFFF
    rp[:300, 1] &= 0x0FFFFFFF
    a = P.unpack_proofs(rp.view(P.PACKED_PROOF).reshape(-1)); b = oracle.packed_unpack_proofs(rp)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.gpu
@pytest.mark.parametrize("algo", ["table", "arith"])
def test_packed_prove_verify_equal_the_oracle(gpu_ctx, oracle, algo):
    """pbh_prove_packed / pbh_verify_packed / pbh_prove_verify_packed: pack(Plonk::prove(unpack(in))) and Plonk::verify of the
    decoded proofs, against the oracle, on both input distributions, across the staged-chunk boundary and on ragged sizes."""
    import pbh_b200 as P
    ctx = gpu_ctx[algo]
    for n, dist in ((1, 1), (255, 0), (70000 + 13, 0), ((1 << 18) + 4097, 1)):
        w, r, c, u, _ = oracle.generate_inputs(n, seed=77, dist=dist, threads=8)
        po, so = oracle.prove_batch(w, r, c, threads=8)
        vo = oracle.verify_batch(po, c, u, threads=8, want_gt=False)
        vo = vo[0] if isinstance(vo, tuple) else vo
        pin = P.pack_witness(w, r, c, u)
        out = ctx.prove_packed(pin)
        assert np.array_equal(_words(out), oracle.packed_pack_proofs(po, so)), (algo, n, dist)
        res = ctx.verify_packed(out, P.pack_chal_u(c, u))
        assert np.array_equal(res, vo), (algo, n, dist)
        out2, res2 = ctx.prove_verify_packed(pin)
        assert np.array_equal(out2, out) and np.array_equal(res2, vo)
    # the reference's own end-to-end vector (src/pbh/mod.rs:44-124) through the packed calls
    w = np.array([[3, 4, 5, 9, 3, 4, 5, 16, 9, 16, 8, 8]], np.uint8).T.copy()
    r = np.array([[7, 4, 11, 12, 16, 2, 14, 11, 7]], np.uint8).T.copy()
    c = np.array([[15, 12, 13, 5, 12]], np.uint8).T.copy()
    u = np.array([4], np.uint8)
    out = ctx.prove_packed(P.pack_witness(w, r, c, u))
    proof, status = P.unpack_proofs(out)
    d = oracle.decode_proof(proof)
    assert int(status[0]) == 0 and (d["a_s"], d["w_z_omega_s"], d["r_z"]) == ((91, 66), (65, 98), 15)
    assert ctx.verify_packed(out, P.pack_chal_u(c, u)).tolist() == [1]


@pytest.mark.gpu
def test_packed_verify_of_arbitrary_words(gpu_ctx, oracle):
    """Words outside the format, tampered digits and non-zero status codes: pbh_verify_packed answers what Plonk::verify (the
    oracle) answers for the planes the restatement decodes them to - PBH_VR_BAD_ENCODING / NOT_IN_FIELD / NOT_ON_CURVE included."""
    import pbh_b200 as P
    rng = np.random.default_rng(9)
    n = 40000
    w, r, c, u, _ = oracle.generate_inputs(n, seed=3, dist=1, threads=8)
    po, so = oracle.prove_batch(w, r, c, threads=8)
    words = oracle.packed_pack_proofs(po, so).copy()
    cu = oracle.packed_pack_chal_u(c, u).copy()
    k = n // 8
    words[0 * k:1 * k, 0] ^= rng.integers(0, 2**32, k, dtype=np.uint64).astype("<u4")          # other points (still on the curve)
    words[1 * k:2 * k, 1] = rng.integers(0, 2**32, k, dtype=np.uint64).astype("<u4")           # ninth digit mostly >= 102
    words[2 * k:3 * k, 2] = rng.integers(0, 2**29, k, dtype=np.uint64).astype("<u4")           # other evaluations, top digit up to 22
    words[3 * k:4 * k, 2] |= rng.integers(1, 8, k, dtype=np.uint64).astype("<u4") << 29        # non-zero status codes
    cu[4 * k:5 * k] = rng.integers(0, 2**32, k, dtype=np.uint64).astype("<u4")                 # challenge words outside the format
    cu[5 * k:6 * k] = rng.integers(0, 17**6, k, dtype=np.uint64).astype("<u4")                 # other challenges (z in H among them)
    planes, _ = oracle.packed_unpack_proofs(words)
    ch, uu = oracle.packed_unpack_chal_u(cu)
    want = oracle.verify_batch(planes, ch, uu, threads=8, want_gt=False)
    want = want[0] if isinstance(want, tuple) else want
    for algo in ("table", "arith"):
        got = gpu_ctx[algo].verify_packed(words.view(P.PACKED_PROOF).reshape(-1), cu)
        assert np.array_equal(got, want), algo
    assert len(set(want.tolist())) >= 5


@pytest.mark.gpu
def test_packed_lanes_and_device_conversions(gpu_ctx, oracle):
    """pbh_prove_packed_async / pbh_verify_packed_async on two lanes over page-locked memory, pageable fallback, and the device
    conversions pbh_unpack_witness_dev / pbh_pack_proof_dev / pbh_unpack_proof_dev against the host codec."""
    import torch
    import pbh_b200 as P
    ctx = gpu_ctx["table"]
    n = 90000 + 5
    sets, expect = [], []
    for lane in range(2):
        w, r, c, u, _ = oracle.generate_inputs(n, first_index=lane * n, seed=21, dist=lane, threads=8)
        po, so = oracle.prove_batch(w, r, c, threads=8)
        vo = oracle.verify_batch(po, c, u, threads=8, want_gt=False)
        expect.append((oracle.packed_pack_proofs(po, so), vo[0] if isinstance(vo, tuple) else vo))
        b = dict(pin=ctx.host_alloc_as(n, P.PACKED_WITNESS), out=ctx.host_alloc_as(n, P.PACKED_PROOF), cu=ctx.host_alloc_as(n, "<u4"),
                 res=ctx.host_alloc_as(n, np.uint8))
        b["pin"][...] = P.pack_witness(w, r, c, u); b["cu"][...] = P.pack_chal_u(c, u)
        b["out"].view(np.uint8)[...] = 0xEE; b["res"][...] = 0xEE
        sets.append((b, (w, r, c, u)))
    for rep in range(3):
        for lane, (b, _) in enumerate(sets):
            ctx.lane_sync(lane)
            ctx.prove_packed_async(lane, b["pin"], b["out"])
            ctx.verify_packed_async(lane, b["out"], b["cu"], b["res"])
    ctx.sync()
    for (b, _), (pw, vo) in zip(sets, expect):
        assert np.array_equal(_words(b["out"]), pw) and np.array_equal(b["res"], vo)
    # the fused call on the lanes: same proofs, same verdicts
    for lane, (b, _) in enumerate(sets):
        b["out"].view(np.uint8)[...] = 0xEE; b["res"][...] = 0xEE
        ctx.prove_verify_packed_async(lane, b["pin"], b["out"], b["res"])
    ctx.sync()
    for (b, _), (pw, vo) in zip(sets, expect):
        assert np.array_equal(_words(b["out"]), pw) and np.array_equal(b["res"], vo)
    # PBH_OPT_PROOF_RESIDENT: the calls above verified the device-resident copy of the proofs (same buffer, same lane, no sync in
    # between).  After a synchronisation the caller may change its buffer, and the verifier must see the change: swap proofs
    # between items, verify on the lane again, compare with the synchronous call on a pageable copy of the tampered buffer.
    b0 = sets[0][0]
    tampered = np.array(b0["out"])
    tampered[:5000] = tampered[5000:10000]
    b0["out"][...] = tampered
    ctx.verify_packed_async(0, b0["out"], b0["cu"], b0["res"])
    ctx.lane_sync(0)
    want_t = ctx.verify_packed(tampered, np.array(b0["cu"]))
    assert np.array_equal(b0["res"], want_t) and not np.array_equal(want_t, expect[0][1])
    # prove, then verify a DIFFERENT buffer on the same lane: uploaded, not taken from the device
    other = ctx.host_alloc_as(n, P.PACKED_PROOF)
    other[...] = tampered
    ctx.prove_packed_async(0, b0["pin"], b0["out"])
    ctx.verify_packed_async(0, other, b0["cu"], b0["res"])
    ctx.lane_sync(0)
    assert np.array_equal(b0["res"], want_t) and np.array_equal(_words(b0["out"]), expect[0][0])
    # the option off gives the same bytes as the resident path
    ctx.set_option(P.OPT_PROOF_RESIDENT, 0)
    ctx.prove_packed_async(0, b0["pin"], b0["out"])
    ctx.verify_packed_async(0, b0["out"], b0["cu"], b0["res"])
    ctx.lane_sync(0)
    ctx.set_option(P.OPT_PROOF_RESIDENT, 1)
    assert np.array_equal(b0["res"], expect[0][1]) and np.array_equal(_words(b0["out"]), expect[0][0])
    ctx.host_free(other)
    # pageable arrays through the lane entry points: synchronous, same bytes
    b, (w, r, c, u) = sets[1]
    out = np.zeros(n, P.PACKED_PROOF); res = np.zeros(n, np.uint8)
    ctx.prove_packed_async(2, np.array(b["pin"]), out)
    ctx.verify_packed_async(2, out, np.array(b["cu"]), res)
    assert np.array_equal(out, b["out"]) and np.array_equal(res, b["res"])
    with pytest.raises(P.PbhError):
        ctx.prove_packed_async(9, np.array(b["pin"]), out)
    # device conversions
    pin_dev = torch.from_numpy(np.array(b["pin"]).view(np.uint8)).cuda()
    dw, dr, dc, du = ctx.unpack_witness_dev(pin_dev, n)
    ctx.sync()
    assert np.array_equal(dw.cpu().numpy(), w) and np.array_equal(dr.cpu().numpy(), r) and np.array_equal(dc.cpu().numpy(), c) and np.array_equal(du.cpu().numpy(), u)
    p_dev, s_dev = ctx.prove_batch(dw, dr, dc)
    packed_dev = ctx.pack_proof_dev(p_dev, s_dev)
    ctx.sync()
    assert np.array_equal(packed_dev.cpu().numpy().view(P.PACKED_PROOF), b["out"])
    cu_dev = torch.from_numpy(np.array(b["cu"]).view(np.uint8)).cuda()
    up, us, uc, uu = ctx.unpack_proof_dev(packed_dev, cu_dev, n)
    ctx.sync()
    hp, hs = P.unpack_proofs(np.array(b["out"]))
    assert np.array_equal(up.cpu().numpy(), hp) and np.array_equal(us.cpu().numpy(), hs)
    assert np.array_equal(uc.cpu().numpy(), c) and np.array_equal(uu.cpu().numpy(), u)
    for bb, _ in sets:
        for arr in bb.values():
            ctx.host_free(arr)
