"""The product's per-item routines (prove_one / verify_one / G1 / pairing of plonk-by-fingers_b200/csrc/*.cuh and the
host setup), compiled for the host by tests/hostemul, against the oracle.  This is the CPU-side guard for the
kernel logic; the same comparisons run on the GPU through the C ABI in test_gpu_parity.py."""
import numpy as np
import pytest


def test_reduction_constants_exhaustive(hostemul):
    """mod17 / mod101 / mod102 shift reciprocals are exact over their whole documented ranges."""
    assert hostemul.check_reductions() == 0


@pytest.mark.parametrize("algo", (0, 1, 2, 3, 4))
@pytest.mark.parametrize("dist", (0, 1))
def test_prove_verify_logic_matches_oracle(hostemul, oracle, algo, dist):
    """algo 0 = ARITH (int32), 1 = TABLE (int32), 2 = TABLE on the FP32 pipes, 3 = ARITH commitments on the FP32 core,
    4 = TABLE on the FP32 pipes with the compile-time constants of the reference's own circuit."""
    circ = oracle.pbh_test_circuit()
    n = 25000
    w, r, c, u, _ = oracle.generate_inputs(n, seed=4242 + dist, dist=dist, threads=8)
    po, so = oracle.prove_batch(w, r, c, threads=8)
    pe, se = hostemul.prove(circ, w, r, c, algo)
    assert np.array_equal(se, so) and np.array_equal(pe, po)
    vo, go = oracle.verify_batch(po, c, u, threads=8)
    ve, ge = hostemul.verify(circ, po, c, u, algo)     # algo 2: TABLE with the FP32-pipe scalar path
    assert np.array_equal(ve, vo) and np.array_equal(ge, go)


def test_f32_bounds(hostemul):
    """Worst-case magnitude propagation through the FP32 prover core: every intermediate is an exactly representable
    integer (< 2^24) and every reduction input is within red17's validated range (<= 2^23), for any circuit constants
    and any SRS length; red17 itself is checked against x mod 17 for every integer |x| <= 2^23."""
    rc, max_exact, max_red = hostemul.f32_bounds()
    assert rc == 0 and max_exact < 2**24 and max_red <= 2**23, (rc, max_exact, max_red)
    assert hostemul.check_red17_f32() == 0


@pytest.mark.parametrize("algo", (0, 1, 2, 3, 4))
def test_zero_blinder_corner_cases(hostemul, oracle, algo):
    """Blinders and challenges drawn from {0, 1, 16}: short polynomials, the Q1 / Q5 / Q15 length logic."""
    rng = np.random.default_rng(17)
    n = 30000
    w = oracle.generate_inputs(n, seed=5, dist=0, threads=8)[0]
    r = rng.choice(np.array([0, 0, 0, 1, 16, 5], dtype=np.uint8), size=(9, n))
    c = rng.choice(np.array([0, 1, 16, 3, 7], dtype=np.uint8), size=(5, n))
    po, so = oracle.prove_batch(w, r, c, threads=8)
    pe, se = hostemul.prove(oracle.pbh_test_circuit(), w, r, c, algo)
    assert np.array_equal(se, so) and np.array_equal(pe, po)
    assert (so == 3).sum() > 0 and (so == 4).sum() > 0      # the rare remainder class is exercised


def test_setup_matches_oracle(hostemul, oracle):
    circ = oracle.pbh_test_circuit()
    for s in range(101):
        for srs_n in (6, 9):
            try:
                og1s, og2, oc = oracle.setup(circuit=circ, s=s, srs_n=srs_n)
                o_ok = True
            except ArithmeticError:
                o_ok = False
            rc, g1s, g2, c = hostemul.setup(circ, s=s, srs_n=srs_n)
            assert (rc == 0) == o_ok, (s, srs_n, rc)
            if o_ok:
                assert np.array_equal(g1s, og1s) and np.array_equal(g2, og2) and np.array_equal(c, oc)


def test_group_law_exhaustive(hostemul, oracle):
    """All 102 x 102 additions and a 102 x 101 scalar-multiple table of the kernels' g1_add / g1_smul."""
    pts = [(x, y, 0) for x in range(101) for y in range(101) if (y * y - x * x * x - 3) % 101 == 0] + [(0, 0, 1)]
    P = np.array(pts, dtype=np.uint8)
    a = np.repeat(P, 102, axis=0); b = np.tile(P, (102, 1))
    exp = oracle.g1_add_batch(np.ascontiguousarray(np.concatenate([a, b], axis=1).T))
    for i in range(0, 102 * 102, 7):
        rc, got = hostemul.g1_add(tuple(a[i]), tuple(b[i]))
        assert rc == 0 and list(got) == exp[:, i].tolist()
    for p in pts:
        for k in (0, 1, 2, 3, 16, 17, 18, 33, 51, 100):
            e = oracle.g1_mul(None if p[2] else p[:2], k)
            e = (0, 0, 1) if e is None else (e[0], e[1], 0)
            assert hostemul.g1_smul(p, k) == e


def test_pairing_exhaustive(hostemul, oracle):
    pts = [(x, y, 0) for x in range(101) for y in range(101) if (y * y - x * x * x - 3) % 101 == 0] + [(0, 0, 1), (1, 2, 1)]
    for q in [(36, 31), (90, 82), (10, 16), (5, 77), (0, 0)]:
        for p in pts:
            e, m = hostemul.pairing(p, q)
            po = None if (p[2] and p[0] == 0) else p
            assert e == oracle.pairing(po, q) and m == oracle.miller(po, q), (p, q)
    for a in range(101):
        for b in range(0, 101, 3):
            assert hostemul.gt_final_exp((a, b)) == oracle.gt_pow((a, b), 600)


def test_q1_inside_the_fp32_core(hostemul, oracle):
    """The SubAssign quirk Q1 (src/poly.rs:192-203) is handled inside the FP32 core since round 2 (no integer fallback):
    SURVEY.md section 9's Q1 tuple, and a batch with zeros forced into a third of the blinders and a fifth of the alphas
    (thousands of status-3 items, every length of t1 + t2), for the three FP32 instantiations."""
    circ = oracle.pbh_test_circuit()
    x, y, z = 4, 4, 7
    w = np.array([[x, y, z, x * x % 17, x, y, z, y * y % 17, x * x % 17, y * y % 17, z * z % 17, z * z % 17]], dtype=np.uint8).T.copy()
    r = np.array([[0, 0, 10, 0, 7, 0, 0, 0, 16]], dtype=np.uint8).T.copy()
    c = np.array([[7, 4, 2, 16, 16]], dtype=np.uint8).T.copy()
    assert oracle.prove_batch(w, r, c)[1].tolist() == [3]
    rng = np.random.default_rng(5)
    n = 120000
    wo, ro, co, uo, _ = oracle.generate_inputs(n, seed=99, dist=0, threads=8)
    ro = ro.copy(); ro[rng.random((9, n)) < 0.35] = 0
    co = co.copy(); co[0, rng.random(n) < 0.2] = 0
    po, so = oracle.prove_batch(wo, ro, co, threads=8)
    assert np.bincount(so, minlength=6)[3] > 500
    for algo in (2, 3, 4):
        assert hostemul.prove(circ, w, r, c, algo)[1].tolist() == [3]
        pe, se = hostemul.prove(circ, wo, ro, co, algo)
        assert np.array_equal(se, so) and np.array_equal(pe, po), algo


def test_half2_reduction_constants():
    """The packed-half reduction of the polynomial sweeps (pbh_kernels.cuh h2_red17): q = fma(x, fp16(1/17), 1536) - 1536,
    r = fma(q, -17, x), emulated with one rounding per fused operation, is THE centred residue for every integer
    |x| <= 2048 (the sweeps stay below 255 * 8 = 2040); the constants are the bit patterns the kernel uses."""
    c17 = np.array([0x2B88], dtype=np.uint16).view(np.float16)[0]
    assert np.array([0x6600, 0xCC40, 0x4C40, 0x6400], dtype=np.uint16).view(np.float16).tolist() == [1536.0, -17.0, 17.0, 1024.0]
    assert c17 == np.float16(1 / 17)
    for v in range(-2048, 2049):
        t = np.float16(np.float64(v) * np.float64(c17) + 1536.0)       # the fused multiply-add rounds once
        q = np.float64(t) - 1536.0
        r = float(np.float16(q * -17.0 + v))
        m = v % 17
        assert r == (m - 17 if m > 8 else m), v
    # the floor reduction of the FP32 polynomial product (f_floor_mod17_biased): canonical residue for 0 <= x < 2^20
    x = np.arange(0, 1 << 20, dtype=np.float64)
    xm8 = (x - 8.0).astype(np.float32)
    t = (xm8.astype(np.float64) * np.float64(np.float32(0.058823529411764705)) + 12582912.0).astype(np.float32)
    q = t - np.float32(12582912.0)
    biased = ((q.astype(np.float64) * -17.0 + xm8.astype(np.float64)).astype(np.float32).astype(np.float64) + 8388616.0).astype(np.float32)
    assert np.array_equal(biased.view(np.uint32) & 0xFF, (np.arange(0, 1 << 20) % 17).astype(np.uint32))
    assert np.array_equal(biased.view(np.uint32) >> 8, np.full(1 << 20, 0x4B0000, np.uint32))


def test_fp32_curve_arithmetic_exhaustive(hostemul, oracle):
    """pbh_g1f.cuh, the exact FP32 arithmetic of PBH_ALGO_ARITH: ALL 102 x 102 sums through the unified slope
    (x1^2 + x1 x2 + x2^2) / (y1 + y2) against the reference's case analysis (src/pbh/g1.rs:119-144), all 102 x 101 multiples
    (:146-168), multiples of off-curve points (they live on y^2 = x^3 + b' and the formulas never use b), the pairing and the
    Miller value of every point against several G2 arguments (src/pbh/pairing.rs:12-47), pow(600) on ALL of F_101^2
    (src/pbh/gt.rs:33-59), and the reduction mod 101 on its whole validity range."""
    assert hostemul.check_red101_f32() == 0
    pts = [(x, y, 0) for x in range(101) for y in range(101) if (y * y - x * x * x - 3) % 101 == 0] + [(0, 0, 1)]
    P = np.array(pts, dtype=np.uint8)
    a = np.repeat(P, 102, axis=0); b = np.tile(P, (102, 1))
    exp = oracle.g1_add_batch(np.ascontiguousarray(np.concatenate([a, b], axis=1).T))
    for i in range(102 * 102):
        assert list(hostemul.g1f_add(tuple(a[i]), tuple(b[i]))) == exp[:, i].tolist(), (a[i], b[i])
    sm = np.ascontiguousarray(np.concatenate([np.repeat(P, 101, axis=0), np.tile(np.arange(101, dtype=np.uint8), 102)[:, None]], axis=1).T)
    exp = oracle.g1_smul_batch(sm)
    for i in range(sm.shape[1]):
        assert list(hostemul.g1f_smul(tuple(sm[:3, i]), int(sm[3, i]))) == exp[:, i].tolist(), sm[:, i]
    rng = np.random.default_rng(12)
    r4 = rng.integers(0, 101, size=(4, 20000), dtype=np.uint8); r4[2] = rng.integers(0, 2, size=20000)      # mostly off-curve
    exp = oracle.g1_smul_batch(r4)
    for i in range(r4.shape[1]):
        assert list(hostemul.g1f_smul(tuple(r4[:3, i]), int(r4[3, i]))) == exp[:, i].tolist(), r4[:, i]
    for q in [(36, 31), (90, 82), (10, 16), (5, 77), (0, 0), (100, 100)]:
        for p in pts + [(1, 2, 1), (48, 0, 1)]:
            e, m = hostemul.pairingf(p, q)
            po = None if (p[2] and p[0] == 0) else p
            assert e == oracle.pairing(po, q) and m == oracle.miller(po, q), (p, q)
    for x in range(101):
        for y in range(101):
            assert hostemul.gtf_final_exp((x, y)) == oracle.gt_pow((x, y), 600)


@pytest.mark.parametrize("algo", (1, 2, 3))
def test_fuzz_arbitrary_bytes(hostemul, oracle, algo):
    """Any byte value in any input plane: same classes and bytes as the oracle (CPU guard of the GPU fuzz test)."""
    rng = np.random.default_rng(321)
    n = 40000
    circ = oracle.pbh_test_circuit()
    wo, ro, co, uo, _ = oracle.generate_inputs(n, seed=56, dist=1, threads=8)
    po, _ = oracle.prove_batch(wo, ro, co, threads=8)

    def corrupt(a):
        a = a.copy()
        mask = rng.random(a.shape) < 0.02
        a[mask] = rng.integers(0, 256, size=int(mask.sum()), dtype=np.uint8)
        return a

    w2, r2, c2 = corrupt(wo), corrupt(ro), corrupt(co)
    pe, se = oracle.prove_batch(w2, r2, c2, threads=8)
    p, s = hostemul.prove(circ, w2, r2, c2, algo)
    assert np.array_equal(s, se) and np.array_equal(p, pe)
    p3, c3, u3 = corrupt(po), corrupt(co), corrupt(uo)
    ve, ge = oracle.verify_batch(p3, c3, u3, threads=8)
    v, g = hostemul.verify(circ, p3, c3, u3, algo)
    assert np.array_equal(v, ve) and np.array_equal(g, ge)


@pytest.mark.parametrize("randomise", ("qc0", "full"))
def test_randomised_circuits_and_srs(hostemul, oracle, randomise):
    """Random selectors and random copy constraints (generally not a permutation, so the grand product does not close and
    the quotient remainder is non-zero without any quirk), constant and zero witnesses that satisfy them, several SRS
    secrets and lengths: the run-time-constant instantiations of the prover and the verifier against the oracle."""
    rng = np.random.default_rng(7 if randomise == "qc0" else 8)
    n = 3000
    seen = set()
    for case in (dict(s=2, srs_n=6), dict(s=7, srs_n=9), dict(s=2, srs_n=4)):
        circ = oracle.pbh_test_circuit()
        for name in ("q_l", "q_r", "q_o", "q_m", "q_c"):
            vals = rng.integers(0, 17, size=4)
            if name == "q_c" and randomise == "qc0":
                vals[:] = 0                    # then the zero witness satisfies the circuit and the deeper sites are reached
            for i in range(4):
                getattr(circ, name)[i] = int(vals[i])
        for name in ("c_a", "c_b", "c_c"):
            ws = rng.integers(0, 3, size=4); idx = rng.integers(1, 5, size=4)
            for i in range(4):
                getattr(circ, name + "_wire")[i] = int(ws[i]); getattr(circ, name + "_index")[i] = int(idx[i])
        wit = rng.integers(0, 17, size=(12, n), dtype=np.uint8)
        wit[:, : n // 2] = wit[:, :1]          # constant witnesses satisfy copy constraints more often
        wit[:, : n // 4] = 0                   # the zero witness satisfies any circuit with q_c = 0
        rnd = rng.integers(0, 17, size=(9, n), dtype=np.uint8)
        chal = rng.integers(0, 17, size=(5, n), dtype=np.uint8)
        u = rng.integers(0, 17, size=n, dtype=np.uint8)
        po, so = oracle.prove_batch(wit, rnd, chal, circuit=circ, threads=8, **case)
        vo, go = oracle.verify_batch(po, chal, u, circuit=circ, threads=8, **case)
        seen |= set(np.unique(so).tolist())
        for algo in (0, 1, 2, 3):
            pe, se = hostemul.prove(circ, wit, rnd, chal, algo, **case)
            assert np.array_equal(se, so) and np.array_equal(pe, po), (case, algo)
            if algo <= 2:
                ve, ge = hostemul.verify(circ, po, chal, u, algo, **case)
                assert np.array_equal(ve, vo) and np.array_equal(ge, go), (case, algo)
    if randomise == "qc0":
        assert 3 in seen          # a remainder that no quirk explains: the copy constraints are not a permutation


@pytest.mark.parametrize("algo", (0, 1, 2))
def test_point_decode_exhaustive(hostemul, oracle, algo):
    """Every byte pair (x, y), with and without the infinity flag, in the place of one proof point: the verifier's
    point decode (table lookup with y folded into 0..50, or in_curve arithmetic) agrees with the oracle on the verdict
    class and on the pairing values."""
    w, r, c, u, _ = oracle.generate_inputs(64, seed=3, dist=1, threads=4)
    p, s = oracle.prove_batch(w, r, c, threads=4)
    ok = np.nonzero(oracle.verify_batch(p, c, u, threads=4, want_gt=False) == 1)[0]
    base = int(ok[0])
    xs, ys = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    for k, flag in ((0, 0), (7, 0), (8, 1), (3, 1)):
        n = 65536
        proof = np.repeat(p[:, base:base + 1], n, axis=1)
        proof[2 * k] = xs.reshape(-1); proof[2 * k + 1] = ys.reshape(-1)
        if flag:
            if k < 8: proof[18] |= np.uint8(1 << k)
            else: proof[19] |= np.uint8(1)
        chal = np.repeat(c[:, base:base + 1], n, axis=1); uu = np.repeat(u[base:base + 1], n)
        ro, go = oracle.verify_batch(proof, chal, uu, threads=8)
        re, ge = hostemul.verify(oracle.pbh_test_circuit(), proof, chal, uu, algo)
        assert np.array_equal(re, ro) and np.array_equal(ge, go), (k, flag)
        assert (ro == 0x20).sum() == 65536 - 101 * 101 and (ro == 0x02).sum() > 0


def test_zz_reduction_argument_ranges(hostemul, oracle):
    """Runs last in this file: every argument the mod 17 / 101 / 102 reductions received during the group-law, pairing,
    prover and verifier tests above (plus an adversarial verifier batch here) stayed inside the range over which their
    shift reciprocals are exact (checked exhaustively by test_reduction_constants_exhaustive)."""
    rng = np.random.default_rng(5)
    n = 4000
    proof = rng.integers(0, 101, size=(27, n), dtype=np.uint8)
    proof[18] = rng.integers(0, 256, size=n); proof[19] = rng.integers(0, 2, size=n); proof[20:] = rng.integers(0, 17, size=(7, n))
    chal = rng.integers(0, 17, size=(5, n), dtype=np.uint8); u = rng.integers(0, 17, size=n, dtype=np.uint8)
    ro, go = oracle.verify_batch(proof, chal, u, threads=8)
    re, ge = hostemul.verify(oracle.pbh_test_circuit(), proof, chal, u, 0)
    assert np.array_equal(re, ro) and np.array_equal(ge, go)
    m17, m101, m102 = hostemul.mod_max_arguments()
    assert 0 < m17 < 69631 and 0 < m101 < 103000 and 0 < m102 < 104000, (m17, m101, m102)
