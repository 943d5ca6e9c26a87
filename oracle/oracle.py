"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The oracle is the C++ restatement of the reference (oracle/pbh_oracle.hpp).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
Batches are numpy uint8 arrays of shape (planes, n), C-contiguous (pitch = n), in the wire layout of
include/pbh_b200.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)
i64p = C.POINTER(C.c_int64)
szp = C.POINTER(C.c_size_t)


def build():
    """Compile liboracle.so with the committed Makefile (g++ only)."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("pbh_oracle.hpp", "oracle_capi.cpp"))
        if not os.path.exists(path) or os.path.getmtime(path) < src_m:
            build()
        _LIB = C.CDLL(path)
        _LIB.oracle_digest.restype = C.c_uint64
    return _LIB


def _p(a):
    return a.ctypes.data_as(u8p)


class Circuit(C.Structure):
    """Layout-identical to pbh_circuit of include/pbh_b200.h."""
    _fields_ = [(name, C.c_uint8 * 4) for name in (
        "q_l", "q_r", "q_o", "q_m", "q_c", "c_a_wire", "c_a_index", "c_b_wire", "c_b_index", "c_c_wire", "c_c_index")]


def pbh_test_circuit():
    """The circuit of src/pbh/mod.rs:56-67: three mul gates, one add gate, and its copy constraints."""
    c = Circuit()
    m1 = 16  # -1 in F_17
    c.q_l[:] = [0, 0, 0, 1]
    c.q_r[:] = [0, 0, 0, 1]
    c.q_o[:] = [m1, m1, m1, m1]
    c.q_m[:] = [1, 1, 1, 0]
    c.q_c[:] = [0, 0, 0, 0]
    A, B, Cc = 0, 1, 2
    c.c_a_wire[:] = [B, B, B, Cc]; c.c_a_index[:] = [1, 2, 3, 1]
    c.c_b_wire[:] = [A, A, A, Cc]; c.c_b_index[:] = [1, 2, 3, 2]
    c.c_c_wire[:] = [A, B, Cc, Cc]; c.c_c_index[:] = [4, 4, 4, 3]
    return c


DEFAULT_SETUP = dict(s=2, srs_n=6, omega_pows=4)


# ---- fine-grained ops -------------------------------------------------------------------------
def field_op(M, op, a, b=0):
    ops = dict(add=0, sub=1, mul=2, div=3, neg=4, pow=5, inv=6)
    out = C.c_uint64()
    rc = lib().oracle_field_op(C.c_uint64(M), ops[op], C.c_uint64(a), C.c_uint64(b), C.byref(out))
    assert rc >= 0
    return None if rc == 1 else out.value


def poly_op(M, op, a, b=(0,)):
    ops = dict(add=0, sub=1, mul=2, div=3, eval=4, z=5, normalize=6, lagrange=7, scale=8)
    a = np.asarray(a, dtype=np.int64); b = np.asarray(b, dtype=np.int64)
    cap = len(a) + len(b) + 4
    o1 = np.zeros(cap, dtype=np.uint64); o2 = np.zeros(cap, dtype=np.uint64)
    l1 = C.c_size_t(); l2 = C.c_size_t()
    rc = lib().oracle_poly_op(C.c_uint64(M), ops[op], a.ctypes.data_as(i64p), C.c_size_t(len(a)),
                              b.ctypes.data_as(i64p), C.c_size_t(len(b)), o1.ctypes.data_as(u64p), C.byref(l1),
                              o2.ctypes.data_as(u64p), C.byref(l2))
    if rc == 1:
        raise ArithmeticError("reference would panic")
    assert rc == 0
    r1 = [int(x) for x in o1[:l1.value]]
    if op == "div":
        return r1, [int(x) for x in o2[:l2.value]]
    if op == "eval":
        return r1[0]
    return r1


def matrix_op(M, op, a, ashape, b=(0,), bshape=(1, 1)):
    ops = dict(mul=0, inv=1, add=2, mul_poly=3)
    a = np.asarray(a, dtype=np.uint64); b = np.asarray(b, dtype=np.uint64)
    out = np.zeros(max(len(a), len(b), ashape[0] * bshape[1]) + 4, dtype=np.uint64)
    om = C.c_size_t(); on = C.c_size_t()
    rc = lib().oracle_matrix_op(C.c_uint64(M), ops[op], a.ctypes.data_as(u64p), C.c_size_t(ashape[0]),
                                C.c_size_t(ashape[1]), b.ctypes.data_as(u64p), C.c_size_t(bshape[0]),
                                C.c_size_t(bshape[1]), out.ctypes.data_as(u64p), C.byref(om), C.byref(on))
    if rc == 1:
        raise ArithmeticError("reference would panic")
    assert rc == 0
    return [int(x) for x in out[:om.value * on.value]], (om.value, on.value)


def fft(M, omega, size, which, inverse, values):
    v = np.asarray(values, dtype=np.uint64)
    out = np.zeros(len(v) + 4, dtype=np.uint64)
    n = C.c_size_t()
    rc = lib().oracle_fft(C.c_uint64(M), C.c_uint64(omega), C.c_size_t(size), dict(vandermonde=0, cooley_tukey=1)[which],
                          int(inverse), v.ctypes.data_as(u64p), C.c_size_t(len(v)), out.ctypes.data_as(u64p), C.byref(n))
    if rc == 1:
        raise ArithmeticError("reference would panic")
    assert rc == 0
    return [int(x) for x in out[:n.value]]


def mul_ntt(M, omega, size, which, a, b):
    a = np.asarray(a, dtype=np.uint64); b = np.asarray(b, dtype=np.uint64)
    out = np.zeros(len(a) + len(b) + 4, dtype=np.uint64)
    n = C.c_size_t()
    rc = lib().oracle_mul_ntt(C.c_uint64(M), C.c_uint64(omega), C.c_size_t(size), dict(vandermonde=0, cooley_tukey=1)[which],
                              a.ctypes.data_as(u64p), C.c_size_t(len(a)), b.ctypes.data_as(u64p), C.c_size_t(len(b)),
                              out.ctypes.data_as(u64p), C.byref(n))
    if rc == 1:
        raise ArithmeticError("reference would panic")
    assert rc == 0
    return [int(x) for x in out[:n.value]]


def _arr(vals):
    return (C.c_uint8 * len(vals))(*vals)


def _g1(p):
    """(x, y) or (x, y, inf) or None (identity)."""
    if p is None:
        return _arr([0, 0, 1])
    return _arr([p[0], p[1], p[2] if len(p) > 2 else 0])


def _g1_out(o):
    return None if (o[2] and o[0] == 0 and o[1] == 0) else ((o[0], o[1]) if not o[2] else (o[0], o[1], 1))


def g1_add(p, q):
    o = _arr([0, 0, 0])
    if lib().oracle_g1_add(_g1(p), _g1(q), o):
        raise ArithmeticError("cannot add")
    return _g1_out(o)


def g1_mul(p, k):
    o = _arr([0, 0, 0])
    if lib().oracle_g1_mul(_g1(p), C.c_uint8(k), o):
        raise ArithmeticError("g1 mul panic")
    return _g1_out(o)


def g1_neg(p):
    o = _arr([0, 0, 0])
    lib().oracle_g1_neg(_g1(p), o)
    return _g1_out(o)


def g1_in_curve(p):
    return bool(lib().oracle_g1_in_curve(_g1(p)))


def g2_add(p, q):
    o = _arr([0, 0])
    if lib().oracle_g2_add(_arr(list(p)), _arr(list(q)), o):
        raise ArithmeticError("g2 add panic")
    return (o[0], o[1])


def g2_mul(p, k):
    o = _arr([0, 0])
    if lib().oracle_g2_mul(_arr(list(p)), C.c_uint8(k), o):
        raise ArithmeticError("g2 mul panic")
    return (o[0], o[1])


def gt_mul(p, q):
    o = _arr([0, 0]); lib().oracle_gt_mul(_arr(list(p)), _arr(list(q)), o); return (o[0], o[1])


def gt_pow(p, n):
    o = _arr([0, 0]); lib().oracle_gt_pow(_arr(list(p)), C.c_uint64(n), o); return (o[0], o[1])


def gt_neg(p):
    o = _arr([0, 0]); lib().oracle_gt_neg(_arr(list(p)), o); return (o[0], o[1])


def pairing(p, q):
    o = _arr([0, 0])
    if lib().oracle_pairing(_g1(p), _arr(list(q)), o):
        raise ArithmeticError("pairing panic")
    return (o[0], o[1])


def miller(p, q):
    o = _arr([0, 0])
    if lib().oracle_miller(_g1(p), _arr(list(q)), o):
        raise ArithmeticError("pairing panic")
    return (o[0], o[1])


# ---- setup / batches ----------------------------------------------------------------------------
def setup(circuit=None, s=2, srs_n=6, omega_pows=4):
    circuit = circuit or pbh_test_circuit()
    g1s = np.zeros(3 * (srs_n + 1), dtype=np.uint8); g2 = np.zeros(4, dtype=np.uint8); consts = np.zeros(24, dtype=np.uint8)
    rc = lib().oracle_setup(C.byref(circuit), C.c_uint8(s), C.c_uint32(srs_n), C.c_uint8(omega_pows), _p(g1s), _p(g2), _p(consts))
    if rc:
        raise ArithmeticError("setup would panic in the reference")
    return g1s.reshape(-1, 3), g2, consts.reshape(8, 3)


def prove_batch(wit, rand, chal, circuit=None, s=2, srs_n=6, omega_pows=4, threads=1):
    circuit = circuit or pbh_test_circuit()
    wit = np.ascontiguousarray(wit, dtype=np.uint8); rand = np.ascontiguousarray(rand, dtype=np.uint8)
    chal = np.ascontiguousarray(chal, dtype=np.uint8)
    n = wit.shape[1]
    assert wit.shape == (12, n) and rand.shape == (9, n) and chal.shape == (5, n)
    proof = np.zeros((27, n), dtype=np.uint8); status = np.zeros(n, dtype=np.uint8)
    rc = lib().oracle_prove_batch(C.byref(circuit), C.c_uint8(s), C.c_uint32(srs_n), C.c_uint8(omega_pows), C.c_size_t(n),
                                  _p(wit), C.c_size_t(n), _p(rand), C.c_size_t(n), _p(chal), C.c_size_t(n), _p(proof),
                                  C.c_size_t(n), _p(status), int(threads))
    if rc:
        raise ArithmeticError("setup would panic in the reference")
    return proof, status


# ---- Fiat-Shamir variants (include/pbh_b200.h, "Fiat-Shamir transcript") --------------------------------
def sha256(data):
    data = np.frombuffer(bytes(data), dtype=np.uint8)
    out = np.zeros(32, dtype=np.uint8)
    lib().oracle_sha256(_p(data) if data.size else None, C.c_size_t(data.size), _p(out))
    return out.tobytes()


def fs_seed(circuit=None, s=2, srs_n=6, omega_pows=4):
    circuit = circuit or pbh_test_circuit()
    out = np.zeros(32, dtype=np.uint8)
    if lib().oracle_fs_seed(C.byref(circuit), C.c_uint8(s), C.c_uint32(srs_n), C.c_uint8(omega_pows), _p(out)):
        raise ArithmeticError("setup would panic in the reference")
    return out.tobytes()


def prove_fs_batch(wit, rand, circuit=None, s=2, srs_n=6, omega_pows=4, threads=1, partial=False):
    """-> proof (27,n), status (n,), chal (6,n) = alpha beta gamma z v u derived from the transcript."""
    circuit = circuit or pbh_test_circuit()
    wit = np.ascontiguousarray(wit, dtype=np.uint8); rand = np.ascontiguousarray(rand, dtype=np.uint8)
    n = wit.shape[1]
    assert wit.shape == (12, n) and rand.shape == (9, n)
    proof = np.zeros((27, n), dtype=np.uint8); status = np.zeros(n, dtype=np.uint8); chal = np.zeros((6, n), dtype=np.uint8)
    rc = lib().oracle_prove_fs_batch(C.byref(circuit), C.c_uint8(s), C.c_uint32(srs_n), C.c_uint8(omega_pows), C.c_size_t(n),
                                     _p(wit), C.c_size_t(n), _p(rand), C.c_size_t(n), _p(proof), C.c_size_t(n), _p(status),
                                     _p(chal), C.c_size_t(n), int(bool(partial)), int(threads))
    if rc:
        raise ArithmeticError("setup would panic in the reference")
    return proof, status, chal


def verify_fs_batch(proof, circuit=None, s=2, srs_n=6, omega_pows=4, threads=1, want_gt=False):
    """-> result (n,), chal (6,n) [, gt (4,n)]"""
    circuit = circuit or pbh_test_circuit()
    proof = np.ascontiguousarray(proof, dtype=np.uint8)
    n = proof.shape[1]
    assert proof.shape == (27, n)
    result = np.zeros(n, dtype=np.uint8); chal = np.zeros((6, n), dtype=np.uint8); gt = np.zeros((4, n), dtype=np.uint8)
    rc = lib().oracle_verify_fs_batch(C.byref(circuit), C.c_uint8(s), C.c_uint32(srs_n), C.c_uint8(omega_pows), C.c_size_t(n),
                                      _p(proof), C.c_size_t(n), _p(result), _p(chal), C.c_size_t(n),
                                      _p(gt) if want_gt else None, C.c_size_t(n), int(threads))
    if rc:
        raise ArithmeticError("setup would panic in the reference")
    return (result, chal, gt) if want_gt else (result, chal)


def verify_batch(proof, chal, u, circuit=None, s=2, srs_n=6, omega_pows=4, threads=1, want_gt=True):
    circuit = circuit or pbh_test_circuit()
    proof = np.ascontiguousarray(proof, dtype=np.uint8); chal = np.ascontiguousarray(chal, dtype=np.uint8)
    u = np.ascontiguousarray(u, dtype=np.uint8)
    n = proof.shape[1]
    assert proof.shape == (27, n) and chal.shape == (5, n) and u.shape == (n,)
    result = np.zeros(n, dtype=np.uint8); gt = np.zeros((4, n), dtype=np.uint8)
    rc = lib().oracle_verify_batch(C.byref(circuit), C.c_uint8(s), C.c_uint32(srs_n), C.c_uint8(omega_pows), C.c_size_t(n),
                                   _p(proof), C.c_size_t(n), _p(chal), C.c_size_t(n), _p(u), _p(result),
                                   _p(gt) if want_gt else None, C.c_size_t(n), int(threads))
    if rc:
        raise ArithmeticError("setup would panic in the reference")
    return (result, gt) if want_gt else result


TRACE_NAMES = ("f_a f_b f_c q_m q_l q_r q_o q_c s1 s2 s3 l1 a b c acc_x z z_omega numerator t r w_z w_z_omega").split()


def prove_trace(w, r, c, circuit=None, s=2, srs_n=6, omega_pows=4):
    circuit = circuit or pbh_test_circuit()
    out = np.zeros(1024, dtype=np.uint8); ln = C.c_size_t()
    status = lib().oracle_prove_trace(C.byref(circuit), C.c_uint8(s), C.c_uint32(srs_n), C.c_uint8(omega_pows),
                                      _arr(list(w)), _arr(list(r)), _arr(list(c)), _p(out), C.c_size_t(len(out)), C.byref(ln))
    assert status >= 0
    polys, pos = {}, 0
    for name in TRACE_NAMES:
        k = int(out[pos]); polys[name] = [int(x) for x in out[pos + 1:pos + 1 + k]]; pos += 1 + k
    return status, polys


def generate_inputs(n, first_index=0, seed=0xB200, dist=1, circuit=None, s=2, srs_n=6, omega_pows=4, threads=1):
    circuit = circuit or pbh_test_circuit()
    wit = np.zeros((12, n), dtype=np.uint8); rand = np.zeros((9, n), dtype=np.uint8); chal = np.zeros((5, n), dtype=np.uint8)
    u = np.zeros(n, dtype=np.uint8); attempt = np.zeros(n, dtype=np.uint8)
    rc = lib().oracle_generate_inputs(C.byref(circuit), C.c_uint8(s), C.c_uint32(srs_n), C.c_uint8(omega_pows), C.c_size_t(n),
                                      C.c_uint64(first_index), C.c_uint64(seed), int(dist), _p(wit), C.c_size_t(n), _p(rand),
                                      C.c_size_t(n), _p(chal), C.c_size_t(n), _p(u), _p(attempt), int(threads))
    assert rc == 0
    return wit, rand, chal, u, attempt


def _planes(fn, nin, nout, arr, *pre, skip_n=False):
    arr = np.ascontiguousarray(arr, dtype=np.uint8)
    n = arr.shape[1]
    assert arr.shape[0] == nin
    out = np.zeros((nout, n), dtype=np.uint8)
    rc = fn(*pre, *(() if skip_n else (C.c_size_t(n),)), _p(arr), C.c_size_t(n), _p(out), C.c_size_t(n))
    assert rc == 0
    return out


def intt4_batch(evals): return _planes(lib().oracle_intt4_batch, 4, 4, evals)
def ntt4_batch(coeffs): return _planes(lib().oracle_ntt4_batch, 4, 4, coeffs)
def g1_smul_batch(arr): return _planes(lib().oracle_g1_smul_batch, 4, 3, arr)
def g1_add_batch(arr): return _planes(lib().oracle_g1_add_batch, 6, 3, arr)
def pairing_batch(arr): return _planes(lib().oracle_pairing_batch, 5, 2, arr)
def kzg_commit_batch(arr, s=2, srs_n=6): return _planes(lib().oracle_kzg_commit_batch, 7, 3, arr, C.c_uint8(s), C.c_uint32(srs_n))


def poly_mul_batch(a, b):
    a = np.ascontiguousarray(a, dtype=np.uint8); b = np.ascontiguousarray(b, dtype=np.uint8)
    la, n = a.shape; lb = b.shape[0]
    out = np.zeros((la + lb - 1, n), dtype=np.uint8)
    rc = lib().oracle_poly_mul_batch(C.c_size_t(n), C.c_uint32(la), C.c_uint32(lb), _p(a), C.c_size_t(n), _p(b), C.c_size_t(n),
                                     _p(out), C.c_size_t(n))
    assert rc == 0
    return out


def poly_add_batch(a, b, subtract=False):
    a = np.ascontiguousarray(a, dtype=np.uint8); b = np.ascontiguousarray(b, dtype=np.uint8)
    ln, n = a.shape
    out = np.zeros((ln, n), dtype=np.uint8)
    rc = lib().oracle_poly_add_batch(C.c_size_t(n), C.c_uint32(ln), int(subtract), _p(a), C.c_size_t(n), _p(b), C.c_size_t(n),
                                     _p(out), C.c_size_t(n))
    assert rc == 0
    return out


def poly_div_zh_batch(p):
    p = np.ascontiguousarray(p, dtype=np.uint8)
    n = p.shape[1]
    q = np.zeros((18, n), dtype=np.uint8); r = np.zeros((4, n), dtype=np.uint8)
    rc = lib().oracle_poly_div_zh_batch(C.c_size_t(n), _p(p), C.c_size_t(n), _p(q), C.c_size_t(n), _p(r), C.c_size_t(n))
    assert rc == 0
    return q, r


def poly_scale_batch(arr):
    """arr: len coefficient planes + 1 scalar plane -> len planes (src/poly.rs:220-228)."""
    return _planes(lib().oracle_poly_scale_batch, arr.shape[0], arr.shape[0] - 1, arr, C.c_size_t(arr.shape[1]), C.c_uint32(arr.shape[0] - 1), skip_n=True)


def poly_eval_batch(arr):
    """arr: len coefficient planes + 1 plane of points -> (n,) evaluations (src/poly.rs:71-79)."""
    arr = np.ascontiguousarray(arr, dtype=np.uint8)
    n = arr.shape[1]
    out = np.zeros(n, dtype=np.uint8)
    rc = lib().oracle_poly_eval_batch(C.c_size_t(n), C.c_uint32(arr.shape[0] - 1), _p(arr), C.c_size_t(n), _p(out))
    assert rc == 0
    return out


def poly_div_linear_batch(arr):
    """arr: len coefficient planes + 1 plane c -> len planes: quotient of p / (x - c) (len-1 planes), then the remainder."""
    return _planes(lib().oracle_poly_div_linear_batch, arr.shape[0], arr.shape[0] - 1, arr, C.c_size_t(arr.shape[1]), C.c_uint32(arr.shape[0] - 1), skip_n=True)


def mul_ntt_batch(a, b, modulus, omega):
    """mul_ntt (src/fft.rs:109-132) over CooleyTurkey: a (la, n), b (lb, n) uint16 -> (la + lb, n)."""
    a = np.ascontiguousarray(a, dtype=np.uint16); b = np.ascontiguousarray(b, dtype=np.uint16)
    la, n = a.shape; lb = b.shape[0]
    out = np.zeros((la + lb, n), dtype=np.uint16)
    u16p = C.POINTER(C.c_uint16)
    rc = lib().oracle_mul_ntt_batch(C.c_size_t(n), C.c_uint64(modulus), C.c_uint64(omega), C.c_uint32(la + lb), C.c_uint32(la), C.c_uint32(lb),
                                    a.ctypes.data_as(u16p), C.c_size_t(n), b.ctypes.data_as(u16p), C.c_size_t(n),
                                    out.ctypes.data_as(u16p), C.c_size_t(n))
    if rc == 1:
        raise ArithmeticError("reference would panic")
    assert rc == 0
    return out


def digest(data, first_index=0):
    data = np.ascontiguousarray(data, dtype=np.uint8)
    if data.ndim == 1:
        data = data.reshape(1, -1)
    planes, n = data.shape
    return int(lib().oracle_digest(C.c_size_t(n), C.c_uint64(first_index), C.c_uint32(planes), _p(data), C.c_size_t(n)))


def pack_verdicts(result):
    result = np.ascontiguousarray(result, dtype=np.uint8)
    out = np.zeros((len(result) + 7) // 8, dtype=np.uint8)
    lib().oracle_pack_verdicts(C.c_size_t(len(result)), _p(result), _p(out))
    return out


def hardware_threads():
    return int(lib().oracle_hardware_threads())


# ---- helpers to build / decode single proofs ----------------------------------------------------
POINT_NAMES = "a_s b_s c_s z_s t_lo_s t_mid_s t_hi_s w_z_s w_z_omega_s".split()
EVAL_NAMES = "a_z b_z c_z s_sigma_1_z s_sigma_2_z r_z z_omega_z".split()


def decode_proof(proof, i=0):
    """Column i of a (27, n) proof batch -> dict of points ((x, y) or None for the identity) and evals."""
    d = {}
    for k, name in enumerate(POINT_NAMES):
        inf = (proof[18, i] >> k) & 1 if k < 8 else proof[19, i] & 1
        x, y = int(proof[2 * k, i]), int(proof[2 * k + 1, i])
        d[name] = (None if (x == 0 and y == 0) else (x, y, 1)) if inf else (x, y)
    for k, name in enumerate(EVAL_NAMES):
        d[name] = int(proof[20 + k, i])
    return d


def encode_proof(d):
    col = np.zeros(27, dtype=np.uint8)
    for k, name in enumerate(POINT_NAMES):
        p = d[name]
        if p is None:
            p = (0, 0, 1)
        col[2 * k], col[2 * k + 1] = p[0], p[1]
        if len(p) > 2 and p[2]:
            if k < 8:
                col[18] |= 1 << k
            else:
                col[19] |= 1
    for k, name in enumerate(EVAL_NAMES):
        col[20 + k] = d[name]
    return col


# ---- packed wire format (include/pbh_b200.h "packed wire format"): an independent restatement of the specification in numpy /
# Python integers, written from the header's text, against which the product's codec (device kernels and host helpers) is
# checked.  The reference has no serialisation of its own (src/plonk.rs:61), so this format is "parity unpinned by the
# reference": what IS pinned is that packed calls equal the byte-plane calls on the decoded values.
def packed_point_codes():
    """code -> (x, y) for the 101 finite points of y^2 = x^3 + 3 over F_101 in increasing (x, y) order; code 0 is the identity."""
    pts = [(0, 0)]
    for x in range(101):
        for y in range(101):
            if (y * y - x * x * x - 3) % 101 == 0:
                pts.append((x, y))
    assert len(pts) == 102
    return pts


def _digits(word, base, count):
    """`count` digits of `word`, least significant first; the last one takes whatever is left and saturates at 255."""
    out = []
    for _ in range(count - 1):
        out.append(word % base)
        word //= base
    out.append(min(word, 255))
    return out


def packed_pack_witness(wit, rand, chal, u):
    vals = np.concatenate([wit, rand, chal, np.asarray(u).reshape(1, -1)], axis=0).astype(np.uint64)      # (27, n)
    assert vals.shape[0] == 27 and int(vals.max(initial=0)) < 17
    w = np.zeros((vals.shape[1], 4), dtype=np.uint64)
    for j, (lo, hi) in enumerate(((0, 7), (7, 14), (14, 21), (21, 27))):
        for k in range(hi - 1, lo - 1, -1):
            w[:, j] = w[:, j] * 17 + vals[k]
    return w.astype("<u4")


def packed_unpack_witness(words):
    """(n, 4) uint32 -> (wit, rand, chal, u)"""
    words = np.asarray(words).reshape(-1, 4)
    n = words.shape[0]
    vals = np.zeros((27, n), dtype=np.uint8)
    for i in range(n):
        d = []
        for j, cnt in enumerate((7, 7, 7, 6)):
            d += _digits(int(words[i, j]), 17, cnt)
        vals[:, i] = d
    return vals[:12].copy(), vals[12:21].copy(), vals[21:26].copy(), vals[26].copy()


def packed_pack_chal_u(chal, u):
    vals = np.concatenate([chal, np.asarray(u).reshape(1, -1)], axis=0).astype(np.uint64)
    w = np.zeros(vals.shape[1], dtype=np.uint64)
    for k in range(5, -1, -1):
        w = w * 17 + vals[k]
    return w.astype("<u4")


def packed_unpack_chal_u(words):
    words = np.asarray(words).reshape(-1)
    vals = np.array([_digits(int(x), 17, 6) for x in words], dtype=np.uint8).T.reshape(6, -1)
    return vals[:5].copy(), vals[5].copy()


_STATUS_CODE = {0: 0, 1: 1, 2: 2, 3: 3, 4: 4, 5: 5, 32: 7}
_CODE_STATUS = {0: 0, 1: 1, 2: 2, 3: 3, 4: 4, 5: 5, 6: 33, 7: 32}


def packed_pack_proofs(proof, status=None):
    """(27, n) planes (+ status) -> (n, 3) uint32 (points_lo, points_hi, evals_status)."""
    pts = packed_point_codes()
    code_of = {p: c for c, p in enumerate(pts) if c > 0}
    n = proof.shape[1]
    out = np.zeros((n, 3), dtype="<u4")
    for i in range(n):
        col = [int(b) for b in proof[:, i]]
        ok = (col[19] & 0xFE) == 0
        codes = []
        for k in range(9):
            x, y = col[2 * k], col[2 * k + 1]
            inf = (col[18] >> k) & 1 if k < 8 else col[19] & 1
            if inf:
                ok = ok and x == 0 and y == 0
                codes.append(0)
            else:
                ok = ok and (x, y) in code_of
                codes.append(code_of.get((x, y), 0))
        evals = col[20:27]
        ok = ok and all(e < 17 for e in evals)
        sc = _STATUS_CODE.get(int(status[i]) if status is not None else 0, 6)
        if sc == 0 and not ok:
            sc = 6
        P = E = 0
        if sc == 0:
            for c in reversed(codes):
                P = P * 102 + c
            for e in reversed(evals):
                E = E * 17 + e
        out[i] = (P & 0xFFFFFFFF, P >> 32, E | (sc << 29))
    return out


def packed_unpack_proofs(words):
    """(n, 3) uint32 -> ((27, n) proof planes, status)"""
    pts = packed_point_codes()
    words = np.asarray(words).reshape(-1, 3)
    n = words.shape[0]
    proof = np.zeros((27, n), dtype=np.uint8)
    status = np.zeros(n, dtype=np.uint8)
    for i in range(n):
        lo, hi, es = (int(x) for x in words[i])
        sc = es >> 29
        status[i] = _CODE_STATUS[sc]
        if sc != 0:
            continue
        P = lo | (hi << 32)
        flags = 0
        for k in range(9):
            d = P % 102 if k < 8 else P
            P //= 102
            if d == 0:
                flags |= 1 << k
            x, y = pts[d] if d < 102 else (0, 0)
            proof[2 * k, i], proof[2 * k + 1, i] = x, y
        proof[18, i], proof[19, i] = flags & 0xFF, flags >> 8
        proof[20:27, i] = _digits(es & ((1 << 29) - 1), 17, 7)
    return proof, status
