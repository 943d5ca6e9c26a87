"""Differential test of the two independent restatements of the reference (C++ oracle vs pure-Python pyref):
the only defence for behaviour the reference's own tests do not pin."""
import numpy as np
import pytest

import pyref


def _circuit_pair(oracle, rng, randomise):
    oc = oracle.pbh_test_circuit()
    pc = {k: list(v) for k, v in pyref.PBH_CIRCUIT.items()}
    if randomise:
        for name in ("q_l", "q_r", "q_o", "q_m", "q_c"):
            vals = [int(x) for x in rng.integers(0, 17, size=4)]
            if name == "q_c" and randomise == "qc0":
                vals = [0] * 4
            pc[name] = vals
            for i in range(4):
                getattr(oc, name)[i] = vals[i]
        for name in ("c_a", "c_b", "c_c"):
            cc = [(int(w), int(i)) for w, i in zip(rng.integers(0, 3, size=4), rng.integers(1, 5, size=4))]
            pc[name] = cc
            for i in range(4):
                getattr(oc, name + "_wire")[i] = cc[i][0]
                getattr(oc, name + "_index")[i] = cc[i][1]
    return oc, pc


def _run_pyref(setup, wit, rnd, chal, u):
    n = wit.shape[1]
    status = np.zeros(n, np.uint8); proof = np.zeros((27, n), np.uint8); result = np.zeros(n, np.uint8); gt = np.zeros((4, n), np.uint8)
    for i in range(n):
        a, b, c = [int(x) for x in wit[0:4, i]], [int(x) for x in wit[4:8, i]], [int(x) for x in wit[8:12, i]]
        try:
            pr = setup.prove(a, b, c, [int(x) for x in rnd[:, i]], [int(x) for x in chal[:, i]])
        except pyref.Panic as e:
            status[i] = e.site
            result[i] = 0xEE
            continue
        for k, p in enumerate(pr["points"]):
            proof[2 * k, i], proof[2 * k + 1, i] = p[0], p[1]
            if p[2]:
                if k < 8: proof[18, i] |= 1 << k
                else: proof[19, i] |= 1
        proof[20:27, i] = pr["evals"]
        try:
            ok, reason, e1, e2 = setup.verify(pr["points"], pr["evals"], [int(x) for x in chal[:, i]], int(u[i]))
            result[i] = {1: 2, 2: 4}.get(reason, 1 if ok else 0)
            if reason == 0:
                gt[:, i] = [e1[0], e1[1], e2[0], e2[1]]
        except pyref.Panic as e:
            result[i] = 0x10 if e.site == 16 else 0x40
    return proof, status, result, gt


@pytest.mark.parametrize("case", [dict(s=2, srs_n=6), dict(s=5, srs_n=6), dict(s=100, srs_n=9), dict(s=2, srs_n=4)])
@pytest.mark.parametrize("randomise", [False, "qc0", "full"])
def test_prove_verify_agree(oracle, case, randomise):
    rng = np.random.default_rng(hash((case["s"], case["srs_n"], str(randomise))) % (1 << 32))
    oc, pc = _circuit_pair(oracle, rng, randomise)
    n = 400
    if not randomise:
        wit, rnd, chal, u, _ = oracle.generate_inputs(n, seed=int(rng.integers(1 << 30)), dist=0)
        w2, r2, c2, u2, _ = oracle.generate_inputs(n // 4, seed=7, dist=1)      # full-path items as well
        wit[:, : n // 4], rnd[:, : n // 4], chal[:, : n // 4], u[: n // 4] = w2, r2, c2, u2
        rnd[:, n // 4: n // 2] *= rng.integers(0, 2, size=(9, n // 4), dtype=np.uint8)   # many zero blinders: Q1, Q5 territory
    else:
        wit = rng.integers(0, 17, size=(12, n), dtype=np.uint8)
        wit[:, : n // 2] = 0
        rnd = rng.integers(0, 17, size=(9, n), dtype=np.uint8); chal = rng.integers(0, 17, size=(5, n), dtype=np.uint8)
        u = rng.integers(0, 17, size=n, dtype=np.uint8)
    setup = pyref.Setup(pc, **case)
    p_py, s_py, v_py, g_py = _run_pyref(setup, wit, rnd, chal, u)
    p_or, s_or = oracle.prove_batch(wit, rnd, chal, circuit=oc, **case)
    assert np.array_equal(s_or, s_py)
    assert np.array_equal(p_or, p_py)
    v_or, g_or = oracle.verify_batch(p_or, chal, u, circuit=oc, **case)
    ok = s_or == 0
    assert np.array_equal(v_or[ok], v_py[ok]) and np.array_equal(g_or[:, ok], g_py[:, ok])


def test_q1_tuple_and_zero_blinders(oracle):
    """(4,4,7) | 0,0,10,0,7,0,0,0,16 | 7,4,2,16,16 panics with a non-zero remainder in both."""
    x, y, z = 4, 4, 7
    a, b, c = [x, y, z, x * x % 17], [x, y, z, y * y % 17], [x * x % 17, y * y % 17, z * z % 17, z * z % 17]
    with pytest.raises(pyref.Panic) as e:
        pyref.Setup(pyref.PBH_CIRCUIT).prove(a, b, c, [0, 0, 10, 0, 7, 0, 0, 0, 16], [7, 4, 2, 16, 16])
    assert e.value.site == 3
    wit = np.array([a + b + c], dtype=np.uint8).T.copy()
    _, st = oracle.prove_batch(wit, np.array([[0, 0, 10, 0, 7, 0, 0, 0, 16]], dtype=np.uint8).T.copy(),
                               np.array([[7, 4, 2, 16, 16]], dtype=np.uint8).T.copy())
    assert st.tolist() == [3]


def test_group_primitives_agree(oracle):
    rng = np.random.default_rng(1)
    pts = [(x, y) for x in range(101) for y in range(101) if (y * y - x ** 3 - 3) % 101 == 0]
    to_py = lambda p: pyref.IDENT if p is None else (p[0], p[1], False)
    to_or = lambda p: None if p[2] else (p[0], p[1])
    for _ in range(1500):
        p = pts[rng.integers(0, 101)] if rng.random() > 0.05 else None
        q = pts[rng.integers(0, 101)] if rng.random() > 0.05 else None
        assert oracle.g1_add(p, q) == to_or(pyref.g1_add(to_py(p), to_py(q)))
        k = int(rng.integers(0, 101))
        assert oracle.g1_mul(p, k) == to_or(pyref.g1_mul(to_py(p), k))
    for p in pts[::3] + [None]:
        for k in (1, 2, 5, 16):
            q = oracle.g2_mul((36, 31), k)
            assert q == pyref.g2_mul((36, 31), k)
            assert oracle.pairing(p, q) == pyref.pairing(to_py(p), q)
    for _ in range(300):
        f = (int(rng.integers(0, 101)), int(rng.integers(0, 101)))
        n = int(rng.integers(0, 700))
        assert oracle.gt_pow(f, n) == pyref.gt_pow(f, n)


def test_sweep_oracles_match_the_polynomial_restatement(oracle):
    """The batch oracles of the scale / eval / divide-by-(x - c) / mul_ntt sweeps are thin loops over the restated Poly
    and fft.rs functions that the reference vectors pin; check the plane packing against the per-polynomial entry points."""
    rng = np.random.default_rng(3)
    arr = rng.integers(0, 17, size=(8, 200), dtype=np.uint8)
    arr[7, :40] = 0
    s, e, d = oracle.poly_scale_batch(arr), oracle.poly_eval_batch(arr), oracle.poly_div_linear_batch(arr)
    pad = lambda v, k: list(v) + [0] * (k - len(v))
    for i in range(200):
        p, c = [int(v) for v in arr[:7, i]], int(arr[7, i])
        assert pad(oracle.poly_op(17, "scale", p, [c]), 7) == s[:, i].tolist()
        assert oracle.poly_op(17, "eval", p, [c]) == e[i]
        q, r = oracle.poly_op(17, "div", p, [(-c) % 17, 1])
        assert pad(q, 6) + pad(r, 1) == d[:, i].tolist()
    a = rng.integers(0, 337, size=(3, 50)).astype(np.uint16); b = rng.integers(0, 337, size=(5, 50)).astype(np.uint16)
    m = oracle.mul_ntt_batch(a, b, 337, 85)
    for i in range(50):
        assert pad(oracle.mul_ntt(337, 85, 8, "cooley_tukey", a[:, i], b[:, i]), 8) == m[:, i].tolist()
    assert oracle.mul_ntt_batch(np.array([[24, 12, 28, 8]], dtype=np.uint16).T, np.array([[4, 26, 29, 23]], dtype=np.uint16).T, 337, 85)[:, 0].tolist() \
        == [96, 335, 109, 312, 285, 202, 184, 0]                     # src/fft.rs:171-183


def test_fiat_shamir_agrees(oracle):
    """The Fiat-Shamir prover of the two restatements: the C++ oracle asks a challenge source inside its templated
    prove; pyref runs its prover as a coroutine driven by hashlib.  Same proofs, statuses and derived challenges."""
    n = 400
    wit, rnd, _, _, _ = oracle.generate_inputs(n, seed=123, dist=0, threads=4)
    rnd[:, :60] = np.random.default_rng(1).choice(np.array([0, 0, 1, 16], dtype=np.uint8), size=(9, 60))
    po, so, co = oracle.prove_fs_batch(wit, rnd, threads=4)
    setup = pyref.Setup(pyref.PBH_CIRCUIT)
    seed = oracle.fs_seed()
    seen = set()
    for i in range(n):
        a, b, c = [int(x) for x in wit[0:4, i]], [int(x) for x in wit[4:8, i]], [int(x) for x in wit[8:12, i]]
        try:
            pr, derived = setup.prove_fs(a, b, c, [int(x) for x in rnd[:, i]], seed)
            st = 0
        except pyref.Panic as e:
            st = e.site
        assert st == so[i], (i, st, so[i])
        seen.add(st)
        if st == 0:
            assert derived == co[:, i].tolist()
            for k, p in enumerate(pr["points"]):
                assert (p[0], p[1]) == (po[2 * k, i], po[2 * k + 1, i])
                assert bool(p[2]) == bool(((po[18, i] >> k) & 1) if k < 8 else (po[19, i] & 1))
            assert pr["evals"] == po[20:27, i].tolist()
    assert seen >= {0, 2, 4, 5}


def test_sweep_oracles_agree_with_pyref(oracle):
    """The same three polynomial sweeps against the second restatement's scale / eval / long division."""
    rng = np.random.default_rng(12)
    arr = rng.integers(0, 17, size=(11, 300), dtype=np.uint8)
    arr[10, :50] = 0
    arr[6:10, 100:200] = 0          # shorter polynomials: the normalised-length paths of both restatements
    s, e, d = oracle.poly_scale_batch(arr), oracle.poly_eval_batch(arr), oracle.poly_div_linear_batch(arr)
    pad = lambda v, k: list(v) + [0] * (k - len(v))
    for i in range(300):
        p, c = pyref.norm([int(v) for v in arr[:10, i]]), int(arr[10, i])
        assert pad(pyref.pscale(p, c), 10) == s[:, i].tolist()
        assert pyref.peval(p, c) == e[i]
        q, r = pyref.pdiv(p, pyref.norm([(-c) % 17, 1]))
        assert pad(q, 9) + pad(r, 1) == d[:, i].tolist()
