"""The C-ABI shared library: it builds for sm_100a, loads, exports every symbol include/pbh_b200.h declares, and
fails loudly without a CUDA device (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pbh_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pbh_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(product_lib):
    import pbh_b200
    declared = _declared_symbols()
    assert len(declared) >= 30
    missing = [s for s in declared if not hasattr(product_lib, s)]
    assert missing == []
    assert sorted(pbh_b200.EXPORTS) == declared


def test_library_targets_sm_100a(product_lib):
    import pbh_b200
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", pbh_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_circuit_struct_layout(product_lib, oracle):
    import pbh_b200
    c = pbh_b200.pbh_test_circuit()
    assert C.sizeof(c) == 44 and bytes(c) == bytes(oracle.pbh_test_circuit())
    assert list(c.q_m) == [1, 1, 1, 0] and list(c.c_c_index) == [4, 4, 4, 3]


def test_no_cpu_fallback(product_lib):
    """Without a GPU, context creation fails with PBH_ERR_NO_DEVICE; with bad parameters it fails before touching CUDA."""
    import pbh_b200
    import torch
    c = pbh_b200.pbh_test_circuit()
    h = C.c_void_p()
    assert product_lib.pbh_ctx_create(C.byref(c), C.c_uint8(2), C.c_uint32(6), C.c_uint8(5), 0, C.byref(h)) == -5   # omega_pows != 4
    assert product_lib.pbh_ctx_create(C.byref(c), C.c_uint8(0), C.c_uint32(6), C.c_uint8(4), 0, C.byref(h)) == -2   # s = 0 panics (Q12)
    assert b"G2" in product_lib.pbh_last_error(None)
    assert product_lib.pbh_ctx_create(C.byref(c), C.c_uint8(17), C.c_uint32(6), C.c_uint8(4), 0, C.byref(h)) == -2  # 17 * G2
    if not torch.cuda.is_available():
        assert product_lib.pbh_ctx_create(C.byref(c), C.c_uint8(2), C.c_uint32(6), C.c_uint8(4), 0, C.byref(h)) == -4
        assert b"no CPU fallback" in product_lib.pbh_last_error(None)
        with pytest.raises(pbh_b200.PbhError):
            pbh_b200.Context()
    # null context / null arguments are reported, never dereferenced
    assert product_lib.pbh_ctx_sync(None) == -1
    assert product_lib.pbh_prove_batch(None, C.c_size_t(1), None, C.c_size_t(1), None, C.c_size_t(1), None, C.c_size_t(1), None,
                                       C.c_size_t(1), None) == -1


def test_product_never_touches_the_oracle():
    """The shipped package must not import, link or execute anything under oracle/ (or tests/)."""
    pkg = os.path.join(ROOT, "plonk-by-fingers_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".rs", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for needle in ("import oracle", "from oracle", "oracle/", "liboracle", "pbh_oracle", "oracle.py", "hostemul", "pyref"):
                    assert needle not in text, (needle, os.path.join(dirpath, f))


def test_expression_front_end_mirrors_the_reference():
    """src/constraints.rs:289-322 (`test_expr`, #[ignore]d upstream because it ends in unimplemented!()): the lowering of
    a*a + b*b - c*c.  The expected trace is the reference's algorithm (:155-196) followed by hand: post-order, one fresh
    variable per operator node, numbered by the size of the variable map at that moment."""
    import pbh_b200
    from pbh_b200 import Constrains, Expression, Gate
    a, b, c = Expression.Var("a"), Expression.Var("b"), Expression.Var("c")
    pitagoras = (a * a) + (b * b) - (c * c)
    assert str(pitagoras) == "(((a*a)+(b*b))-(c*c))"
    variables, gates = {}, []
    r = Constrains.eval_exprs(pitagoras, variables, gates)
    assert r == 7 and variables == {"a": 0, "v1": 1, "b": 2, "v3": 3, "v4": 4, "c": 5, "v6": 6, "v7": 7}
    sel = lambda g: (g.q_l, g.q_r, g.q_o, g.q_m, g.q_c)
    mul, add, sub = sel(Gate.mul_a_b()), sel(Gate.sum_a_b()), sel(Gate.sub_a_b())
    assert [(sel(g), l, rr, o) for g, l, rr, o in gates] == [(mul, 0, 0, 1), (mul, 2, 2, 3), (add, 1, 3, 4), (mul, 5, 5, 6), (sub, 4, 6, 7)]
    assert sub == (1, 1, 1, 0, 0)                       # sic: src/constraints.rs:37-45
    with pytest.raises(pbh_b200.ReferencePanic):        # Expression::Const => unimplemented!()  (:166-168)
        Constrains.eval_exprs(a + Expression.Const(3), {}, [])
    # five gates: more than the four the reference's prove() is hard-wired to, which is why upstream stops here
    with pytest.raises(pbh_b200.PbhError):
        Constrains([g for g, _, _, _ in gates], ([], [], []))
