"""Pins the CPU oracle (oracle/pbh_oracle.hpp) against every known-answer vector of the reference's own unit tests
(tests/golden/reference_vectors.json, transcribed from the cited file:line) and against the derived vectors of
SURVEY.md §9.  The oracle is the checker for the GPU parity tests, so it is pinned first."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_vectors.json")) as f:
    V = json.load(f)


def test_fixture_is_regenerable():
    """The committed JSON equals what the committed generator script holds."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(HERE, "golden", "make_reference_vectors.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    assert json.loads(json.dumps(mk.V)) == V


def test_field_vectors(oracle):
    """src/utils/u64field.rs:238-254"""
    for v in V["field_f101"]:
        op = v["op"]
        if op in ("add", "sub", "pow"):
            assert oracle.field_op(101, op, v["a"], v["b"]) == v["out"]
        elif op == "div":
            assert oracle.field_op(101, "div", v["a"], v["b"]) == v["out"]
        elif op == "div_then_mul":
            q = oracle.field_op(101, "div", v["a"], v["b"])
            assert oracle.field_op(101, "mul", v["b"], q) == v["out"]
        elif op == "neg":
            assert oracle.field_op(101, "neg", v["a"]) == v["out"]
        elif op == "neg_div":
            assert oracle.field_op(101, "neg", oracle.field_op(101, "div", v["a"], v["b"])) == v["out"]
    # inverse is defined for every non-zero element of both fields and 0 has none
    for m in (17, 101):
        assert oracle.field_op(m, "inv", 0) is None
        for a in range(1, m):
            assert oracle.field_op(m, "mul", a, oracle.field_op(m, "inv", a)) == 1


def test_poly_vectors(oracle):
    """src/poly.rs:402-487"""
    P = V["poly_f15485863"]; M = P["modulus"]
    red = lambda c: [x % M for x in c]
    for a, b, out in P["add"]:
        assert oracle.poly_op(M, "add", a, b) == red(out)
    for a, b, out in P["sub"]:
        assert oracle.poly_op(M, "sub", a, b) == red(out)
    for a, b, out in P["mul"]:
        assert oracle.poly_op(M, "mul", a, b) == red(out)
    for n, d in P["div_roundtrip"]:
        q, r = oracle.poly_op(M, "div", n, d)
        back = oracle.poly_op(M, "add", oracle.poly_op(M, "mul", q, d), r)
        assert back == oracle.poly_op(M, "normalize", n)
    xs = [p[0] for p in P["lagrange_points"]]; ys = [p[1] for p in P["lagrange_points"]]
    l = oracle.poly_op(M, "lagrange", xs, ys)
    for x, y in zip(xs, ys):
        assert oracle.poly_op(M, "eval", l, [x]) == y
    for pts, out in P["z"]:
        assert oracle.poly_op(M, "z", pts) == red(out)
    for c, x, y in P["eval"]:
        assert oracle.poly_op(M, "eval", c, [x]) == y
    for c, out in P["normalize"]:
        assert oracle.poly_op(M, "normalize", c) == out


def test_poly_quirks(oracle):
    """Q1 (src/poly.rs:192-203): a longer rhs leaves its tail un-negated; Q15 (:220-228): times zero is [0]."""
    assert oracle.poly_op(17, "sub", [1, 2], [1, 2, 3, 4]) == [0, 0, 3, 4]
    assert oracle.poly_op(17, "sub", [1, 2, 3, 4], [1, 2]) == [0, 0, 3, 4]
    assert oracle.poly_op(17, "sub", [0], [5, 6]) == [12, 6]
    assert oracle.poly_op(17, "scale", [1, 2, 3], [0]) == [0]
    assert oracle.poly_op(17, "scale", [1, 2, 3], [2]) == [2, 4, 6]


def test_matrix_vectors(oracle):
    """src/matrix.rs:203-227"""
    Mx = V["matrix_f104729"]; M = Mx["modulus"]
    out, shape = oracle.matrix_op(M, "add", Mx["add"]["a"], Mx["add"]["shape"], Mx["add"]["b"], Mx["add"]["shape"])
    assert out == Mx["add"]["out"]
    out, shape = oracle.matrix_op(M, "mul", Mx["mul"]["a"], Mx["mul"]["ashape"], Mx["mul"]["b"], Mx["mul"]["bshape"])
    assert out == Mx["mul"]["out"] and shape == (2, 2)
    a, sh = Mx["inv_involution"]["a"], Mx["inv_involution"]["shape"]
    inv, _ = oracle.matrix_op(M, "inv", a, sh)
    assert inv != a
    assert oracle.matrix_op(M, "inv", inv, sh)[0] == a
    # Plonk::new's interpolation matrix over F_17
    h = [1, 4, 16, 13]
    vand = [pow(h[r], c, 17) for r in range(4) for c in range(4)]
    hinv, _ = oracle.matrix_op(17, "inv", vand, (4, 4))
    assert hinv == [13, 13, 13, 13, 13, 16, 4, 1, 13, 4, 13, 4, 13, 1, 4, 16]


def test_fft_vectors(oracle):
    """src/fft.rs:140-183"""
    F = V["fft_f337"]; M, w, n = F["modulus"], F["omega"], F["size"]
    for which in ("vandermonde", "cooley_tukey"):
        freq = oracle.fft(M, w, n, which, False, F["values"])
        assert freq == F["freq"]
        assert oracle.fft(M, w, n, which, True, freq) == F["values"]
    school = oracle.poly_op(M, "mul", F["mul_a"], F["mul_b"])
    ntt = oracle.poly_op(M, "normalize", oracle.mul_ntt(M, w, n, "cooley_tukey", F["mul_a"], F["mul_b"]))
    assert ntt == school == [96, 335, 109, 312, 285, 202, 184]
    # Q14: the Vandermonde variant returns a shorter vector when trailing outputs are zero
    assert len(oracle.fft(M, w, n, "vandermonde", False, [0] * 8)) == 1
    assert len(oracle.fft(M, w, n, "cooley_tukey", False, [0] * 8)) == 8


def test_g1_vectors(oracle):
    """src/pbh/g1.rs:233-260"""
    g = V["g1"]; G = tuple(g["generator"])
    two, four = oracle.g1_add(G, G), None
    four = oracle.g1_add(two, two); eight = oracle.g1_add(four, four); sixteen = oracle.g1_add(eight, eight)
    T = lambda k: tuple(g[k])
    assert oracle.g1_neg(G) == T("neg_g") and two == T("two_g") and oracle.g1_neg(two) == T("neg_two_g")
    assert four == T("four_g") and oracle.g1_neg(four) == T("neg_four_g")
    assert eight == T("eight_g") and oracle.g1_neg(eight) == T("neg_eight_g")
    assert sixteen == T("sixteen_g") and oracle.g1_neg(sixteen) == T("neg_sixteen_g")
    assert oracle.g1_add(two, G) == T("two_g_plus_g") and oracle.g1_add(four, G) == T("four_g_plus_g")
    assert oracle.g1_add(eight, G) == T("eight_g_plus_g")
    assert oracle.g1_add(four, two) == oracle.g1_add(two, four)
    assert oracle.g1_mul(G, 1) == G and oracle.g1_mul(G, 2) == two
    six = G
    for _ in range(5):
        six = oracle.g1_add(six, G)
    assert oracle.g1_mul(G, 6) == six
    # the subgroup table of SURVEY.md §9 and the identity conventions (Q9, Q16)
    table = [None, (1, 2), (68, 74), (26, 45), (65, 98), (12, 32), (32, 42), (91, 35), (18, 49), (18, 52), (91, 66), (32, 59),
             (12, 69), (65, 3), (26, 56), (68, 27), (1, 99), None]
    for k, exp in enumerate(table):
        assert oracle.g1_mul(G, k) == exp
    assert oracle.g1_add(G, oracle.g1_neg(G)) is None and oracle.g1_add(None, G) == G and oracle.g1_add(G, None) == G
    assert not oracle.g1_in_curve(None) and oracle.g1_in_curve((1, 2, 1)) and oracle.g1_in_curve((48, 0))
    assert oracle.g1_add((48, 0), (48, 0)) is None            # order-2 point: doubling is caught by self == -rhs
    assert sum(oracle.g1_in_curve((x, y)) for x in range(101) for y in range(101)) == 101


def test_g2_gt_pairing_vectors(oracle):
    """src/pbh/g2.rs:108-119, src/pbh/gt.rs:88-97, src/pbh/pairing.rs:56-75"""
    g2 = tuple(V["g2"]["generator"])
    assert oracle.g2_add(g2, g2) == tuple(V["g2"]["two_g"])
    two = oracle.g2_add(g2, g2)
    assert oracle.g2_add(two, two) == oracle.g2_add(oracle.g2_add(oracle.g2_add(g2, g2), g2), g2)
    six = g2
    for _ in range(5):
        six = oracle.g2_add(six, g2)
    assert oracle.g2_mul(g2, 6) == six
    with pytest.raises(ArithmeticError):
        oracle.g2_mul(g2, 0)                                   # Q12
    with pytest.raises(ArithmeticError):
        oracle.g2_mul(g2, 17)
    for a, b, out in V["gt"]["mul"]:
        assert oracle.gt_mul(a, b) == tuple(out)
    for a, n, out in V["gt"]["pow"]:
        assert oracle.gt_pow(a, n) == tuple(out)
    base = V["gt"]["frobenius_base"]
    assert oracle.gt_pow(base, 101) == oracle.gt_neg(base)
    assert oracle.gt_pow(base, 102) == oracle.gt_mul(oracle.gt_neg(base), base)
    # bilinearity exactly as the reference tests it
    pv = V["pairing"]
    G = (1, 2)
    p, r, q, a = oracle.g1_mul(G, pv["p_scalar"]), oracle.g1_mul(G, pv["r_scalar"]), oracle.g2_mul(g2, pv["q_scalar"]), pv["a"]
    e = oracle.pairing
    assert e(oracle.g1_mul(p, a), q) == e(p, oracle.g2_mul(q, a))
    assert e(oracle.g1_mul(p, a), q) == oracle.gt_pow(e(p, q), a)
    assert e(oracle.g1_add(p, r), q) == oracle.gt_mul(e(p, q), e(r, q))
    assert e(oracle.g1_mul(p, a), q) == (97, 12)              # SURVEY.md §9 [derived]
    # derived absolute values
    exp = [(7, 28), (97, 89), (38, 6), (31, 96), (93, 25), (59, 52), (26, 97), (2, 94), (2, 7), (26, 4), (59, 49), (93, 76), (31, 5),
           (38, 95), (97, 12), (7, 73)]
    for k in range(1, 17):
        assert e(oracle.g1_mul(G, k), g2) == exp[k - 1]
    assert oracle.miller(G, g2) == (15, 26) and e(None, g2) == (0, 0)
    assert e((6, 44), g2) == (31, 5) and oracle.miller((6, 44), g2) == (6, 88) and e((48, 0), g2) == (0, 0)
    g2s = [(36, 31), (90, 82), (10, 16), (63, 35), (74, 12), (41, 22), (66, 23), (2, 34), (2, 67), (66, 78), (41, 79), (74, 89), (63, 66),
           (10, 85), (90, 19), (36, 70)]
    for k in range(1, 17):
        assert oracle.g2_mul(g2, k) == g2s[k - 1]


def test_plonk_gen_proof(oracle):
    """src/pbh/mod.rs:44-124: the reference's only end-to-end vector, plus the pairing value of README.md:6."""
    g = V["plonk_gen_proof"]
    wit = np.array([g["a"] + g["b"] + [c % 17 for c in g["c"]]], dtype=np.uint8).T.copy()
    rnd = np.array([g["rand"]], dtype=np.uint8).T.copy()
    ch = g["challange"]
    chal = np.array([[ch["alpha"], ch["beta"], ch["gamma"], ch["z"], ch["v"]]], dtype=np.uint8).T.copy()
    proof, status = oracle.prove_batch(wit, rnd, chal, s=g["s"], srs_n=g["srs_n"], omega_pows=g["omega_pows"])
    assert status.tolist() == [0]
    d = oracle.decode_proof(proof)
    for k, v in g["proof"].items():
        assert d[k] == (tuple(v) if isinstance(v, list) else v), k
    res, gt = oracle.verify_batch(proof, chal, np.array(g["verify_rand"], dtype=np.uint8))
    assert bool(res[0] & 1) == g["verify"] and res[0] == 1
    assert gt[:, 0].tolist() == g["pairing_value"] * 2


def test_golden_run_intermediates(oracle):
    """SURVEY.md §9 [derived]: every intermediate polynomial of the golden run."""
    st, tr = oracle.prove_trace([3, 4, 5, 9, 3, 4, 5, 16, 9, 16, 8, 8], [7, 4, 11, 12, 16, 2, 14, 11, 7], [15, 12, 13, 5, 12])
    assert st == 0
    exp = dict(f_a=[1, 13, 3, 3], f_b=[7, 3, 14, 13], f_c=[6, 5, 11, 4], q_m=[5, 16, 13, 1], q_l=[13, 1, 4, 16], q_r=[13, 1, 4, 16],
               q_o=[16], q_c=[0], s1=[7, 13, 10, 6], s2=[4, 0, 13, 1], s3=[6, 7, 3, 14], l1=[13, 13, 13, 13], a=[14, 6, 3, 3, 4, 7],
               b=[12, 9, 14, 13, 12, 11], c=[4, 6, 11, 4, 2, 16], acc_x=[0, 16, 5, 14], z=[10, 5, 8, 14, 7, 11, 14],
               z_omega=[10, 3, 9, 12, 7, 10, 3],
               numerator=[6, 1, 4, 8, 11, 3, 0, 1, 16, 11, 3, 7, 3, 13, 11, 16, 8, 12, 16, 2, 7, 11],
               t=[11, 16, 13, 9, 0, 13, 13, 8, 1, 2, 10, 1, 15, 6, 16, 2, 7, 11], r=[0, 16, 9, 13, 8, 15, 16], w_z=[16, 13, 2, 9, 3, 5],
               w_z_omega=[13, 14, 2, 13, 2, 14])
    for k, v in exp.items():
        assert tr[k] == v, k
    g1s, g2, consts = oracle.setup()
    assert g1s.tolist() == [[1, 2, 0], [68, 74, 0], [65, 98, 0], [18, 49, 0], [1, 99, 0], [68, 27, 0], [65, 3, 0]]
    assert g2.tolist() == [36, 31, 90, 82]
    assert consts.tolist() == [[12, 69, 0], [32, 42, 0], [32, 42, 0], [1, 99, 0], [0, 0, 1], [68, 74, 0], [65, 3, 0], [18, 49, 0]]


def test_status_distribution(oracle):
    """outcome shares on uniform inputs (first failing site wins), within sampling error."""
    n = 40000
    w, r, c, u, _ = oracle.generate_inputs(n, seed=99, dist=0, threads=8)
    proof, status = oracle.prove_batch(w, r, c, threads=8)
    res = oracle.verify_batch(proof, c, u, threads=8, want_gt=False)
    h = np.bincount(status, minlength=6) / n
    assert abs(h[2] - 0.424) < 0.015 and abs(h[4] - 0.149) < 0.01 and abs(h[5] - 0.315) < 0.015 and h[1] == 0 and h[3] < 0.003
    ok = status == 0
    share = lambda code: float(((res == code) & ok).sum()) / n
    assert abs(share(2) - 0.046) < 0.006 and abs(share(0x10) - 0.015) < 0.004
    assert abs(share(0) - 0.019) < 0.005 and abs(share(1) - 0.032) < 0.005
