"""pytest configuration: `gpu` marker, import paths, shared fixtures.

`-m "not gpu"` covers the oracle against the reference's vectors, the host logic, the kernel routines compiled for
the host (tests/hostemul) and the C-ABI surface; `-m gpu` tests are the parity tests proper and call the CUDA
path through the C ABI.  Nothing here reads /root/reference at run time.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "plonk-by-fingers_b200", "python"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def product_lib():
    """libpbh_b200.so, built in-tree if missing (nvcc cross-compiles without a GPU)."""
    import pbh_b200
    if not os.path.exists(pbh_b200.LIB_PATH):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "plonk-by-fingers_b200")], check=True)
    return pbh_b200.load_library()


@pytest.fixture(scope="session")
def hostemul():
    import hostemul_binding
    return hostemul_binding.load()


def _cuda_ok():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_ctx(product_lib):
    """One context per algorithm on cuda:0.  GPU tests FAIL (not skip) when the extension cannot run."""
    import pbh_b200
    assert _cuda_ok(), "gpu tests need a CUDA device"
    ctxs = {algo: pbh_b200.Context(device=0, algo=algo) for algo in ("table", "arith")}
    # "table" runs the prover's arithmetic on the FP32 pipes (default); "table_int" is the same algorithm on int32 IMAD
    ctxs["table_int"] = pbh_b200.Context(device=0, algo="table")
    ctxs["table_int"].set_option(pbh_b200.OPT_PROVER_FP32, 0)
    # per-item curve arithmetic with the int32 prover (the default "arith" context uses the FP32 core for the F_17 work)
    ctxs["arith_int"] = pbh_b200.Context(device=0, algo="arith")
    ctxs["arith_int"].set_option(pbh_b200.OPT_PROVER_FP32, 0)
    # FP32 prover, generic instantiation (run-time circuit constants) even though the circuit is the reference's own
    ctxs["table_generic"] = pbh_b200.Context(device=0, algo="table")
    ctxs["table_generic"].set_option(pbh_b200.OPT_SPECIALISE, 0)
    # the verifier's F_17 scalar work on the FP32 pipes (the default is int32)
    ctxs["table_vf32"] = pbh_b200.Context(device=0, algo="table")
    ctxs["table_vf32"].set_option(pbh_b200.OPT_VERIFIER_FP32, 1)
    # FP32 prover with plain per-thread loads/stores instead of TMA-staged tiles
    ctxs["table_notma"] = pbh_b200.Context(device=0, algo="table")
    ctxs["table_notma"].set_option(pbh_b200.OPT_TMA, 0)
    yield ctxs
    for c in ctxs.values():
        c.close()
