"""ctypes binding of tests/hostemul/libhostemul.so (TEST INFRASTRUCTURE): the product's __host__ __device__
per-item routines compiled with g++, so kernel logic can be checked against the oracle without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostemul")
ROOT = os.path.dirname(os.path.dirname(HERE))
u8p = C.POINTER(C.c_uint8)
_E = None


def load():
    global _E
    if _E is None:
        so = os.path.join(HERE, "libhostemul.so")
        srcs = [os.path.join(HERE, "hostemul.cpp")] + [os.path.join(ROOT, "plonk-by-fingers_b200", "csrc", f) for f in
                                                       ("pbh_arith.cuh", "pbh_prove.cuh", "pbh_prove_f32.cuh", "pbh_verify.cuh", "pbh_setup.hpp", "pbh_sha256.cuh", "pbh_fs.cuh", "pbh_f32.cuh", "pbh_g1f.cuh")]
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-mfma", "-DPBH_RANGE_TRACK", "-o", so, srcs[0]], check=True)
        _E = C.CDLL(so)
        _E.emul_last_error.restype = C.c_char_p
        _E.emul_check_reductions.restype = C.c_uint64
        _E.emul_tables_size.restype = C.c_size_t
    return Emul(_E)


def _p(a):
    return a.ctypes.data_as(u8p)


class Emul:
    def __init__(self, lib):
        self.lib = lib

    def prove(self, circuit, wit, rand, chal, algo, s=2, srs_n=6, omega_pows=4):
        wit, rand, chal = (np.ascontiguousarray(x, dtype=np.uint8) for x in (wit, rand, chal))
        n = wit.shape[1]
        proof = np.zeros((27, n), np.uint8); status = np.zeros(n, np.uint8)
        rc = self.lib.emul_prove_batch(C.byref(circuit), C.c_uint8(s), C.c_uint32(srs_n), C.c_uint8(omega_pows), int(algo),
                                       C.c_size_t(n), _p(wit), _p(rand), _p(chal), _p(proof), _p(status))
        if rc:
            raise RuntimeError((rc, self.lib.emul_last_error().decode()))
        return proof, status

    def verify(self, circuit, proof, chal, u, algo, s=2, srs_n=6, omega_pows=4):
        proof, chal, u = (np.ascontiguousarray(x, dtype=np.uint8) for x in (proof, chal, u))
        n = proof.shape[1]
        res = np.zeros(n, np.uint8); gt = np.zeros((4, n), np.uint8)
        rc = self.lib.emul_verify_batch(C.byref(circuit), C.c_uint8(s), C.c_uint32(srs_n), C.c_uint8(omega_pows), int(algo),
                                        C.c_size_t(n), _p(proof), _p(chal), _p(u), _p(res), _p(gt))
        if rc:
            raise RuntimeError((rc, self.lib.emul_last_error().decode()))
        return res, gt

    # ---- Fiat-Shamir routines ----
    def sha256(self, data):
        data = np.frombuffer(bytes(data), dtype=np.uint8)
        out = np.zeros(32, np.uint8)
        self.lib.emul_sha256(_p(data) if data.size else None, C.c_size_t(data.size), _p(out))
        return out.tobytes()

    def sha256_absorb(self, state, msg):
        st = np.frombuffer(bytes(state), dtype=np.uint8).copy()
        m = np.frombuffer(bytes(msg), dtype=np.uint8)
        self.lib.emul_sha256_absorb(_p(st), _p(m) if m.size else None, int(m.size))
        return st.tobytes()

    def fs_seed(self, circuit, s=2, srs_n=6, omega_pows=4):
        out = np.zeros(32, np.uint8)
        rc = self.lib.emul_fs_seed(C.byref(circuit), C.c_uint8(s), C.c_uint32(srs_n), C.c_uint8(omega_pows), _p(out))
        if rc:
            raise RuntimeError((rc, self.lib.emul_last_error().decode()))
        return out.tobytes()

    def prove_fs(self, circuit, wit, rand, algo, s=2, srs_n=6, omega_pows=4):
        wit, rand = (np.ascontiguousarray(x, dtype=np.uint8) for x in (wit, rand))
        n = wit.shape[1]
        proof = np.zeros((27, n), np.uint8); status = np.zeros(n, np.uint8); chal = np.zeros((6, n), np.uint8)
        rc = self.lib.emul_prove_fs_batch(C.byref(circuit), C.c_uint8(s), C.c_uint32(srs_n), C.c_uint8(omega_pows), int(algo),
                                          C.c_size_t(n), _p(wit), _p(rand), _p(proof), _p(status), _p(chal))
        if rc:
            raise RuntimeError((rc, self.lib.emul_last_error().decode()))
        return proof, status, chal

    def verify_fs(self, circuit, proof, algo, s=2, srs_n=6, omega_pows=4):
        proof = np.ascontiguousarray(proof, dtype=np.uint8)
        n = proof.shape[1]
        res = np.zeros(n, np.uint8); chal = np.zeros((6, n), np.uint8); gt = np.zeros((4, n), np.uint8)
        rc = self.lib.emul_verify_fs_batch(C.byref(circuit), C.c_uint8(s), C.c_uint32(srs_n), C.c_uint8(omega_pows), int(algo),
                                           C.c_size_t(n), _p(proof), _p(res), _p(chal), _p(gt))
        if rc:
            raise RuntimeError((rc, self.lib.emul_last_error().decode()))
        return res, chal, gt

    def setup(self, circuit, s=2, srs_n=6, omega_pows=4):
        g1s = np.zeros(3 * (srs_n + 1), np.uint8); g2 = np.zeros(4, np.uint8); consts = np.zeros(24, np.uint8)
        rc = self.lib.emul_setup(C.byref(circuit), C.c_uint8(s), C.c_uint32(srs_n), C.c_uint8(omega_pows), _p(g1s), _p(g2), _p(consts),
                                 None, C.c_size_t(0))
        return rc, g1s.reshape(-1, 3), g2, consts.reshape(8, 3)

    def g1_add(self, p, q):
        o = (C.c_uint8 * 3)()
        rc = self.lib.emul_g1_add((C.c_uint8 * 3)(*p), (C.c_uint8 * 3)(*q), o)
        return rc, tuple(o)

    def g1_smul(self, p, k):
        o = (C.c_uint8 * 3)()
        self.lib.emul_g1_smul((C.c_uint8 * 3)(*p), C.c_uint8(k), o)
        return tuple(o)

    def pairing(self, p, q):
        o = (C.c_uint8 * 2)(); m = (C.c_uint8 * 2)()
        self.lib.emul_pairing((C.c_uint8 * 3)(*p), (C.c_uint8 * 2)(*q), o, m)
        return tuple(o), tuple(m)

    # ---- the exact FP32 curve arithmetic of pbh_g1f.cuh (PBH_ALGO_ARITH on the device) ----
    def g1f_add(self, p, q):
        o = (C.c_uint8 * 3)()
        self.lib.emul_g1f_add((C.c_uint8 * 3)(*p), (C.c_uint8 * 3)(*q), o)
        return tuple(o)

    def g1f_smul(self, p, k):
        o = (C.c_uint8 * 3)()
        self.lib.emul_g1f_smul((C.c_uint8 * 3)(*p), C.c_uint8(k), o)
        return tuple(o)

    def pairingf(self, p, q):
        o = (C.c_uint8 * 2)(); m = (C.c_uint8 * 2)()
        self.lib.emul_pairingf((C.c_uint8 * 3)(*p), (C.c_uint8 * 2)(*q), o, m)
        return tuple(o), tuple(m)

    def gtf_final_exp(self, f):
        o = (C.c_uint8 * 2)()
        self.lib.emul_gtf_final_exp((C.c_uint8 * 2)(*f), o)
        return tuple(o)

    def check_red101_f32(self):
        self.lib.emul_check_red101_f32.restype = C.c_uint64
        return int(self.lib.emul_check_red101_f32())

    def gt_final_exp(self, f):
        o = (C.c_uint8 * 2)()
        self.lib.emul_gt_final_exp((C.c_uint8 * 2)(*f), o)
        return tuple(o)

    def f32_bounds(self):
        mx = C.c_double(); mr = C.c_double()
        rc = self.lib.emul_f32_bounds(C.byref(mx), C.byref(mr))
        return rc, mx.value, mr.value

    def check_red17_f32(self):
        self.lib.emul_check_red17_f32.restype = C.c_uint64
        return int(self.lib.emul_check_red17_f32())

    def check_reductions(self):
        return int(self.lib.emul_check_reductions())

    def mod_max_arguments(self):
        """largest arguments seen by mod17, mod101, mod102 in this process"""
        self.lib.emul_mod_max_argument.restype = C.c_uint32
        return tuple(int(self.lib.emul_mod_max_argument(i)) for i in range(3))
