"""CPU: the pin of SURVEY 8f row 3 against the reference itself.

tests/golden/ceres_pnp.npz holds 284 problems solved by REAL Ceres 2.0 (the reference's prebuilt libceres.so.2.0.0 + its
unmodified src/uncertainty_pnp.cpp, tests/golden/make_golden_ceres.py).  Checked here, without a GPU:
  * oracle/pnp_oracle.py (the numpy restatement) follows Ceres: same stop reason, same number of iterations, same cost after
    every iteration, same pose (1e-9) -- including descents with up to 20 rejected steps and the 50-iteration cap;
  * the arithmetic core of the CUDA kernel (csrc/pnp_core.cuh compiled as host code) does the same;
  * where oracle/_ref/ceres exists (this container), live Ceres reproduces the fixture bit for bit, and the reference's own
    C entry `uncertainty_pnp` returns the same bits as the instrumented solve.
Iteration bookkeeping: Ceres appends an IterationSummary when an iteration is finalised; an iteration that ends the solve by
the parameter- or function-tolerance test returns before that (trust_region_minimizer.cc), so for those reasons
#summaries == iterations started (summary 0 is the initial evaluation), otherwise #summaries - 1."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "ceres_pnp.npz"))
N = len(G["pn"])


def problem(i):
    pn = int(G["pn"][i])
    return G["pts2d"][i, :pn], G["pts3d"][i, :pn], G["wgt2d"][i, :pn], G["K"][i], G["init_rt"][i]


def expected_iterations(i):
    return int(G["iteration_summaries"][i]) - (0 if G["reason"][i] in (2, 3) else 1)


def followable(i):
    return bool(G["stable"][i]) and G["kind"][i] != "optimum"


def test_fixture_shape():
    assert N == 284 and (G["linear_solver_type_used"] == 3).all()           # DENSE_SCHUR, as uncertainty_pnp.cpp:84 asks
    assert sum(followable(i) for i in range(N)) >= 230
    assert int(G["unsuccessful"][[followable(i) for i in range(N)]].max()) >= 15   # rejected steps are covered
    assert (G["reason"][[followable(i) for i in range(N)]] == 5).sum() >= 10       # and so is the iteration cap
    assert np.array_equal(G["result_rt"], G["entry_rt"], equal_nan=True)   # probe == the reference's own C entry


def test_oracle_follows_ceres():
    import pnp_oracle as po
    worst = 0.0
    for i in range(N):
        with np.errstate(all="ignore"):
            x, info = po.uncertainty_pnp(*problem(i), return_info=True)
        if not followable(i):
            if G["kind"][i] == "optimum":                                  # costs ~1e-20: only the pose is meaningful
                assert np.abs(x - G["result_rt"][i]).max() < 1e-8, i
            continue
        assert info["termination"] == G["reason"][i], i
        assert info["iterations"] == expected_iterations(i), i
        d = np.abs(x - G["result_rt"][i]).max()
        worst = max(worst, d)
        assert d < 1e-9, (i, d)
        assert abs(info["cost"] - G["final_cost"][i]) <= 1e-9 * G["final_cost"][i], i
    assert worst < 1e-9


@pytest.fixture(scope="module")
def host_core():
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libpnp_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-x", "c++",
                           os.path.join(ROOT, "tests", "pnp_host_harness.cpp"), "-o", so])
    lib = ctypes.CDLL(so)
    DP = ctypes.POINTER(ctypes.c_double)

    def solve(uv, p3, W, K, init):
        a = [np.ascontiguousarray(v, np.float64) for v in (uv, p3, W, K, init)]
        res, info = np.empty(6), (ctypes.c_int * 2)()
        lib.pnp_host_solve(*[v.ctypes.data_as(DP) for v in a], res.ctypes.data_as(DP), info, ctypes.c_int(len(uv)),
                           ctypes.c_int(50), ctypes.c_double(1e-6), ctypes.c_double(1e-10), ctypes.c_double(1e-8))
        return res, info[0], info[1]
    return solve


def test_cuda_core_follows_ceres(host_core):
    """csrc/pnp_core.cuh (the code the kernel runs, compiled for the host) against real Ceres."""
    for i in range(N):
        if not followable(i):
            continue
        x, it, code = host_core(*problem(i))
        assert (it, code) == (expected_iterations(i), G["reason"][i]), i
        assert np.abs(x - G["result_rt"][i]).max() < 1e-9, i


def test_live_ceres_reproduces_the_fixture():
    import build_ceres_ref as ceres
    if not ceres.available():
        pytest.skip("oracle/_ref/ceres not built (needs the reference checkout)")
    for i in list(range(0, N, 7)):
        with np.errstate(all="ignore"):
            res, info, tr = ceres.solve(*problem(i))
            ent = ceres.reference_entry(*problem(i))
        assert np.array_equal(res, G["result_rt"][i], equal_nan=True) and np.array_equal(ent, res, equal_nan=True), i
        assert info["reason"] == G["reason"][i] and info["iteration_summaries"] == G["iteration_summaries"][i]
        k = min(len(tr), G["cost_trace"].shape[1])
        assert np.array_equal(tr[:k, 1], G["cost_trace"][i, :k], equal_nan=True)
