"""pyref.py — a SECOND, independently written restatement of the reference (TEST INFRASTRUCTURE).

Pure Python, small cases only.  It exists to differential-test oracle/pbh_oracle.hpp: the reference pins the hot
path with one end-to-end vector, so everything else is pinned by two restatements agreeing.
It follows the reference's data-dependent behaviour (normalised coefficient lists, panic order, quirks Q1-Q17),
written from the reference text, not from the C++ oracle.  Citations are relative to the reference repository.
"""


class Panic(Exception):
    def __init__(self, site, what=""):
        super().__init__(f"{site}: {what}")
        self.site = site


P17, P101 = 17, 101
OMEGA, K1, K2 = 4, 2, 3          # src/pbh/mod.rs:27-29


# ---- src/utils/u64field.rs ---------------------------------------------------------------------
def inv(a, m):
    """Extended Euclid, None when not invertible (:10-25, :52-63)."""
    r0, r1, s0, s1 = a, m, 1, 0
    while r1:
        q = r0 // r1
        r0, r1 = r1, r0 - q * r1
        s0, s1 = s1, s0 - q * s1
    if r0 != 1:
        return None
    return s0 % m


def fdiv(a, b, m):
    i = inv(b, m)
    return None if i is None else (i * a) % m


# ---- src/poly.rs ---------------------------------------------------------------------------------
def norm(c):
    c = list(c)
    while len(c) > 1 and c[-1] == 0:
        c.pop()
    return c or [0]


def padd(a, b, m=P17):
    out = list(a)
    for n in range(max(len(a), len(b))):
        if n >= len(out):
            out.append(b[n])
        elif n < len(b):
            out[n] = (out[n] + b[n]) % m
    return norm(out)


def psub(a, b, m=P17):
    """Q1: the tail of a longer rhs is appended, not negated (:192-203)."""
    out = list(a)
    for n in range(max(len(a), len(b))):
        if n >= len(out):
            out.append(b[n])
        elif n < len(b):
            out[n] = (out[n] - b[n]) % m
    return norm(out)


def pmul(a, b, m=P17):
    out = [0] * (len(a) + len(b))
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            out[i + j] = (out[i + j] + x * y) % m
    return norm(out)


def pscale(a, k, m=P17):
    """Q15 (:220-228)."""
    if k % m == 0:
        return [0]
    return [(x * k) % m for x in a]


def padd_scalar(a, k, m=P17):
    out = list(a)
    out[0] = (out[0] + k) % m
    return norm(out)


def peval(a, x, m=P17):
    xp, y = 1, a[0]
    for i in range(1, len(a)):
        xp = (xp * x) % m
        y = (y + xp * a[i]) % m
    return y


def pdiv(a, b, m=P17):
    """Long division (:230-247)."""
    q, r = [0], list(a)
    while r != [0] and len(r) >= len(b):
        lead = (r[-1] * inv(b[-1], m)) % m
        t = norm([0] * (len(r) - len(b)) + [lead])
        q = padd(q, t, m)
        r = psub(r, pmul(b, t, m), m)
    return norm(q), norm(r)


# ---- src/pbh/g1.rs -------------------------------------------------------------------------------
IDENT = (0, 0, True)


def g1(x, y):
    return (x % P101, y % P101, False)


def g1_neg(p):
    return p if p[2] else (p[0], (-p[1]) % P101, False)


def g1_add(p, q):
    if p[2]:
        return q
    if q[2]:
        return p
    if p == g1_neg(q):
        return IDENT
    if p == q:
        m = fdiv(3 * p[0] * p[0] % P101, 2 * p[1] % P101, P101)
        if m is None:
            raise Panic(64, "g1 double")
        x = (m * m - 2 * p[0]) % P101
        return (x, (m * (3 * p[0] - m * m) - p[1]) % P101, False)
    lam = fdiv((q[1] - p[1]) % P101, (q[0] - p[0]) % P101, P101)
    if lam is None:
        raise Panic(64, "cannot add")
    x = (lam * lam - p[0] - q[0]) % P101
    return (x, (lam * (p[0] - x) - p[1]) % P101, False)


def g1_mul(p, k):
    k %= P101
    if k == 0 or p[2]:
        return IDENT
    res, base = None, p
    while k:
        if k & 1:
            res = base if res is None else g1_add(res, base)
        k >>= 1
        base = g1_add(base, base)
    return res


def in_curve(p):
    return (p[1] * p[1]) % P101 == (p[0] ** 3 + 3) % P101


G = g1(1, 2)


# ---- src/pbh/g2.rs, gt.rs, pairing.rs ------------------------------------------------------------
def g2_add(p, q):
    if p == q:
        m_u = fdiv(3 * p[0] * p[0] % P101, 2 * p[1] % P101, P101)
        if m_u is None:
            raise Panic(64, "g2 double")
        ui = inv((-2) % P101, P101)
        m2 = m_u * m_u * ui % P101
        return ((m2 - 2 * p[0]) % P101, (ui * m_u * (3 * p[0] - m2) - p[1]) % P101)
    lam = fdiv((q[1] - p[1]) % P101, (q[0] - p[0]) % P101, P101)
    if lam is None:
        raise Panic(64, "g2 add")
    a = (lam * lam * -2 - p[0] - q[0]) % P101
    return (a, (lam * (p[0] - a) - p[1]) % P101)


def g2_mul(p, k):
    res, base = None, p
    while k:
        if k & 1:
            res = base if res is None else g2_add(res, base)
        k >>= 1
        base = g2_add(base, base)
    if res is None:
        raise Panic(64, "g2 * 0")
    return res


def gt_mul(p, q):
    return ((p[0] * q[0] - 2 * p[1] * q[1]) % P101, (p[0] * q[1] + p[1] * q[0]) % P101)


def gt_pow(p, n):
    if n >= 101:
        c = gt_pow(p, n // 101)
        acc = (c[0], (-c[1]) % P101)
        n %= 101
    else:
        acc = (1, 0)
    base = p
    while n:
        if n & 1:
            acc = gt_mul(acc, base)
        n >>= 1
        base = gt_mul(base, base)
    return acc


def pairing_f(r, p, q):
    def line(a, b):
        m, n = (b[0] - a[0]) % P101, (b[1] - a[1]) % P101
        return n, (-m) % P101, (m * a[1] - n * a[0]) % P101
    if r == 1:
        return (1, 0)
    if r % 2 == 1:
        r -= 1
        x, y, c = line(g1_mul(p, r), p)
        return gt_mul(pairing_f(r, p, q), ((q[0] * x + c) % P101, q[1] * y % P101))
    r //= 2
    x, y, c = line(g1_mul(p, r), g1_mul(g1_mul(g1_neg(p), r), 2))
    return gt_mul(gt_pow(pairing_f(r, p, q), 2), ((q[0] * x + c) % P101, q[1] * y % P101))


def pairing(p, q):
    return gt_pow(pairing_f(17, p, q), (101 ** 2 - 1) // 17)


# ---- src/plonk.rs ----------------------------------------------------------------------------------
class Setup:
    def __init__(self, circuit, s=2, srs_n=6):
        """circuit: dict with q_l q_r q_o q_m q_c (lists of 4) and c_a c_b c_c (lists of (wire 0..2, index 1..4))."""
        self.c = circuit
        self.g1s = [G]
        sp = s % P101
        for _ in range(srs_n):
            self.g1s.append(g1_mul(G, sp))
            sp = sp * s % P101                      # Q11
        self.g2_1 = (36, 31)
        self.g2_s = g2_mul(self.g2_1, s % P101)
        self.h = [pow(OMEGA, n, P17) for n in range(4)]
        self.k1h = [x * K1 % P17 for x in self.h]
        self.k2h = [x * K2 % P17 for x in self.h]
        # inverse Vandermonde by solving: interpolation == evaluating the unique degree<4 interpolant
        self.zh = [16, 0, 0, 0, 1]

    def interp(self, vals):
        """h_pows_inv * vals (src/plonk.rs:153-160, 177-179) computed as Lagrange interpolation over H."""
        out = [0]
        for j, hj in enumerate(self.h):
            num, den = [1], 1
            for i, hi in enumerate(self.h):
                if i != j:
                    num = pmul(num, [(-hi) % P17, 1])
                    den = den * (hj - hi) % P17
            out = padd(out, pscale(num, vals[j] * inv(den, P17) % P17))
        return norm(out)

    def roots(self, cc):
        return [(self.h, self.k1h, self.k2h)[w][n - 1] for (w, n) in cc]

    def commit(self, poly):
        acc = IDENT
        for n, v in enumerate(poly):
            if n >= len(self.g1s):
                raise Panic(5, "srs oob")
            acc = g1_add(acc, g1_mul(self.g1s[n], v))
        return acc

    def satisfies(self, a, b, c):
        k = self.c
        for n in range(4):
            if (k["q_l"][n] * a[n] + k["q_l"][n] * b[n] + k["q_o"][n] * c[n] + k["q_m"][n] * a[n] * b[n] + k["q_c"][n]) % P17:
                return False                        # Q8
        wires = (a, b, c)
        for n in range(4):
            for vec, cc in ((a, k["c_a"]), (b, k["c_b"]), (c, k["c_c"])):
                w, i = cc[n]
                if vec[n] != wires[w][i - 1]:
                    return False
        return True

    def prove(self, a, b, c, rand, ch):
        """src/plonk.rs:191-466 with the caller's Challange (alpha, beta, gamma, z, v)."""
        alpha, beta, gamma, z, v = ch
        answers = {"beta_gamma": (beta, gamma), "alpha": alpha, "z": z, "v": v, "u": None}
        steps = self._prove_steps(a, b, c, rand)
        try:
            ask, _ = next(steps)
            while True:
                ask, _ = steps.send(answers[ask])
        except StopIteration as done:
            return done.value

    def prove_fs(self, a, b, c, rand, seed):
        """Fiat-Shamir (include/pbh_b200.h): the same prover with each challenge derived from the SHA-256 transcript at
        the moment it is first needed.  Returns (proof dict, [alpha, beta, gamma, z, v, u])."""
        import hashlib
        state = [bytes(seed)]
        derived = [0] * 6

        def absorb(msg):
            state[0] = hashlib.sha256(state[0] + msg).digest()
            return state[0]

        def point(p):
            return bytes([p[0], p[1], 1 if p[2] else 0, 0])

        def squeeze(st, k):
            return int.from_bytes(st[8 * k:8 * k + 8], "big") % P17

        steps = self._prove_steps(a, b, c, rand)
        try:
            ask, data = next(steps)
            while True:
                if ask == "v":
                    st = absorb(bytes(data))
                else:
                    st = absorb(b"".join(point(p) for p in data))
                if ask == "beta_gamma":
                    derived[1], derived[2] = squeeze(st, 0), squeeze(st, 1)
                    answer = (derived[1], derived[2])
                else:
                    idx = {"alpha": 0, "z": 3, "v": 4, "u": 5}[ask]
                    derived[idx] = answer = squeeze(st, 0)
                ask, data = steps.send(answer)
        except StopIteration as done:
            return done.value, derived

    def _prove_steps(self, a, b, c, rand):
        """The prover as a coroutine: yields (which challenge it needs now, the transcript data produced since the last
        one) at the point where the reference first uses that challenge, and returns the proof."""
        k = self.c
        if not self.satisfies(a, b, c):
            raise Panic(1)
        s1v, s2v, s3v = self.roots(k["c_a"]), self.roots(k["c_b"]), self.roots(k["c_c"])
        fa, fb, fc = self.interp(a), self.interp(b), self.interp(c)
        qo, qm, ql, qr, qc = (self.interp(k[n]) for n in ("q_o", "q_m", "q_l", "q_r", "q_c"))
        S1, S2, S3 = self.interp(s1v), self.interp(s2v), self.interp(s3v)
        b1, b2, b3, b4, b5, b6, b7, b8, b9 = rand
        ax = padd(pmul(norm([b2, b1]), self.zh), fa)
        bx = padd(pmul(norm([b4, b3]), self.zh), fb)
        cx = padd(pmul(norm([b6, b5]), self.zh), fc)
        a_s, b_s, c_s = self.commit(ax), self.commit(bx), self.commit(cx)
        beta, gamma = yield ("beta_gamma", [a_s, b_s, c_s])
        acc = [1]
        for i in range(1, 4):
            w = pow(OMEGA, i - 1, P17)
            dend = (a[i - 1] + beta * w + gamma) * (b[i - 1] + beta * K1 * w + gamma) * (c[i - 1] + beta * K2 * w + gamma) % P17
            dsor = ((a[i - 1] + beta * peval(S1, w) + gamma) * (b[i - 1] + beta * peval(S2, w) + gamma) *
                    (c[i - 1] + beta * peval(S3, w) + gamma)) % P17
            d = fdiv(dend, dsor, P17)
            if d is None:
                raise Panic(2)
            acc.append(acc[-1] * d % P17)
        accx = self.interp(acc)
        if peval(accx, pow(OMEGA, 4, P17)) != 1:
            raise Panic(8)
        zx = padd(pmul(norm([b9, b8, b7]), self.zh), accx)
        z_s = self.commit(zx)
        alpha = yield ("alpha", [z_s])
        L1 = self.interp([1, 0, 0, 0])
        t1 = padd(padd(padd(padd(padd(pmul(pmul(ax, bx), qm), pmul(ax, ql)), pmul(bx, qr)), pmul(cx, qo)), [0]), qc)
        A2 = pscale(padd(ax, norm([gamma, beta])), alpha)
        B2 = padd(bx, norm([gamma, beta * K1 % P17]))
        C2 = padd(cx, norm([gamma, beta * K2 % P17]))
        zw = norm([zx[n] * pow(OMEGA, n, P17) % P17 for n in range(len(zx))])
        A3 = pscale(padd_scalar(padd(ax, pscale(S1, beta)), gamma), alpha)
        B3 = padd_scalar(padd(bx, pscale(S2, beta)), gamma)
        C3 = padd_scalar(padd(cx, pscale(S3, beta)), gamma)
        t4 = pmul(pscale(padd(zx, [16]), alpha * alpha % P17), L1)
        t2 = pmul(pmul(pmul(A2, B2), C2), zx)
        t3 = pmul(pmul(pmul(A3, B3), C3), zw)
        numer = padd(psub(padd(t1, t2), t3), t4)
        tx, rem = pdiv(numer, self.zh)
        if rem != [0]:
            raise Panic(3)
        if len(tx) < 18:
            raise Panic(4)
        thi, tmid, tlo = norm(tx[12:18]), norm(tx[6:12]), norm(tx[0:6])
        t_hi_s, t_mid_s, t_lo_s = self.commit(thi), self.commit(tmid), self.commit(tlo)
        z = yield ("z", [t_lo_s, t_mid_s, t_hi_s])
        a_z, b_z, c_z = peval(ax, z), peval(bx, z), peval(cx, z)
        s1z, s2z = peval(S1, z), peval(S2, z)
        t_z = peval(tx, z)
        zwz = peval(zw, z)
        r1 = padd(padd(padd(padd(pscale(pscale(qm, a_z), b_z), pscale(ql, a_z)), pscale(qr, b_z)), pscale(qo, c_z)), qc)
        r2 = pscale(zx, (a_z + beta * z + gamma) * (b_z + beta * K1 * z + gamma) * (c_z + beta * K2 * z + gamma) * alpha % P17)
        r3 = pscale(pmul(zx, pscale(pscale(S3, beta), zwz)), (a_z + beta * s1z + gamma) * (b_z + beta * s2z + gamma) * alpha % P17)  # Q2
        r4 = pscale(pscale(zx, peval(L1, z)), alpha * alpha % P17)
        rx = padd(padd(padd(r1, r2), r3), r4)
        r_z = peval(rx, z)
        v = yield ("v", [a_z, b_z, c_z, s1z, s2z, r_z, zwz])
        wn = padd_scalar(padd(padd(tlo, pscale(tmid, pow(z, 6, P17))), pscale(thi, pow(z, 12, P17))), -t_z)
        for poly, val, e in ((rx, r_z, 1), (ax, a_z, 2), (bx, b_z, 3), (cx, c_z, 4), (S1, s1z, 5), (S2, s2z, 6)):
            wn = padd(wn, pscale(padd_scalar(poly, -val), pow(v, e, P17)))
        wz, rem = pdiv(wn, norm([(-z) % P17, 1]))
        if rem != [0]:
            raise Panic(6)
        wzw, rem = pdiv(padd_scalar(zx, -zwz), norm([(-z) % P17 * OMEGA % P17, 1]))
        if rem != [0]:
            raise Panic(7)
        w_z_s = self.commit(wz)
        w_zw_s = self.commit(wzw)
        yield ("u", [w_z_s, w_zw_s])
        return dict(points=[a_s, b_s, c_s, z_s, t_lo_s, t_mid_s, t_hi_s, w_z_s, w_zw_s], evals=[a_z, b_z, c_z, s1z, s2z, r_z, zwz])

    def verify(self, points, evals, ch, u):
        """Returns (verdict, reason, e1, e2); raises Panic(16) when Z_H(z) = 0.  Evals may be raw bytes >= 17."""
        alpha, beta, gamma, z, v = ch
        k = self.c
        s1v, s2v, s3v = self.roots(k["c_a"]), self.roots(k["c_b"]), self.roots(k["c_c"])
        qm_s, ql_s, qr_s, qo_s, qc_s = (self.commit(self.interp(k[n])) for n in ("q_m", "q_l", "q_r", "q_o", "q_c"))
        s1_s, s2_s, s3_s = self.commit(self.interp(s1v)), self.commit(self.interp(s2v)), self.commit(self.interp(s3v))
        if not all(in_curve(p) for p in points):
            return False, 1, None, None
        if not all(e < P17 for e in evals):
            return False, 2, None, None
        a_s, b_s, c_s, z_s, t_lo_s, t_mid_s, t_hi_s, w_z_s, w_zw_s = points
        a_z, b_z, c_z, s1z, s2z, r_z, zwz = evals
        zh_z = peval(self.zh, z)
        l1_z = peval(self.interp([1, 0, 0, 0]), z)
        t_z = fdiv((r_z - (beta * s1z + gamma + a_z) * (beta * s2z + gamma + b_z) * (c_z + gamma) * zwz - l1_z * alpha * alpha) % P17, zh_z, P17)
        if t_z is None:
            raise Panic(16)
        d1 = g1_add(g1_add(g1_add(g1_add(g1_mul(qm_s, a_z * b_z * v % P17), g1_mul(ql_s, a_z * v % P17)), g1_mul(qr_s, b_z * v % P17)),
                           g1_mul(qo_s, c_z * v % P17)), g1_mul(qc_s, v))
        d2 = g1_mul(z_s, ((a_z + beta * z + gamma) * (b_z + beta * K1 * z + gamma) * (c_z + beta * K2 * z + gamma) * alpha * v +
                          l1_z * alpha * alpha * v + u) % P17)
        d3 = g1_mul(s3_s, (a_z + beta * s1z + gamma) * (b_z + beta * s2z + gamma) * alpha * v * beta * zwz % P17)
        d = g1_add(g1_add(d1, d2), g1_neg(d3))
        f = t_lo_s
        for p, sc in ((t_mid_s, pow(z, 6, P17)), (t_hi_s, pow(z, 12, P17))):
            f = g1_add(f, g1_mul(p, sc))
        f = g1_add(f, d)
        for p, e in ((a_s, 2), (b_s, 3), (c_s, 4), (s1_s, 5), (s2_s, 6)):
            f = g1_add(f, g1_mul(p, pow(v, e, P17)))
        e_s = g1_mul(self.commit([1]), (t_z + v * r_z + pow(v, 2, P17) * a_z + pow(v, 3, P17) * b_z + pow(v, 4, P17) * c_z +
                                        pow(v, 5, P17) * s1z + pow(v, 6, P17) * s2z + u * zwz) % P17)
        e1q1 = g1_add(w_z_s, g1_mul(w_zw_s, u))
        e2q1 = g1_add(g1_add(g1_add(g1_mul(w_z_s, z), g1_mul(w_zw_s, u * z * OMEGA % P17)), f), g1_neg(e_s))
        e1, e2 = pairing(e1q1, self.g2_s), pairing(e2q1, self.g2_1)
        return e1 == e2, 0, e1, e2


PBH_CIRCUIT = dict(q_l=[0, 0, 0, 1], q_r=[0, 0, 0, 1], q_o=[16] * 4, q_m=[1, 1, 1, 0], q_c=[0] * 4,
                   c_a=[(1, 1), (1, 2), (1, 3), (2, 1)], c_b=[(0, 1), (0, 2), (0, 3), (2, 2)], c_c=[(0, 4), (1, 4), (2, 4), (2, 3)])
