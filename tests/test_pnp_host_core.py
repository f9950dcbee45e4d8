"""CPU: the arithmetic core of the CUDA uncertainty-PnP kernel (clean_pvnet_b200/csrc/pnp_core.cuh), compiled as host code
by tests/pnp_host_harness.cpp, against the oracle.  The kernel (csrc/pnp.cu) adds only the warp reduction of the normal
equations around this core, so residuals, Jacobians, the 6x6 solve and every branch of the trust-region state machine are
checked here without a GPU; tests/test_gpu_pnp.py then checks the kernel itself."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from util import pnp_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DP = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def host():
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libpnp_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-x", "c++",
                           os.path.join(ROOT, "tests", "pnp_host_harness.cpp"), "-o", so])
    lib = ctypes.CDLL(so)

    def P(a):
        return a.ctypes.data_as(DP)

    def solve(uv, p3, W, K, init, mi=50, ft=1e-6, gt=1e-10, pt=1e-8):
        uv, p3, W, K, init = [np.ascontiguousarray(a, np.float64) for a in (uv, p3, W, K, init)]
        out, info = np.empty(6), (ctypes.c_int * 2)()
        lib.pnp_host_solve(P(uv), P(p3), P(W), P(K), P(init), P(out), info, ctypes.c_int(len(uv)), ctypes.c_int(mi),
                           ctypes.c_double(ft), ctypes.c_double(gt), ctypes.c_double(pt))
        return out, (info[0], info[1])

    def normal(pose, uv, p3, W, K):
        pose, uv, p3, W, K = [np.ascontiguousarray(a, np.float64) for a in (pose, uv, p3, W, K)]
        H, g, c = np.empty(21), np.empty(6), ctypes.c_double()
        lib.pnp_host_normal(P(pose), P(uv), P(p3), P(W), P(K), ctypes.c_int(len(uv)), P(H), P(g), ctypes.byref(c))
        return H, g, c.value

    solve.normal = normal
    return solve


@pytest.fixture(scope="module")
def po():
    import pnp_oracle
    return pnp_oracle


@pytest.mark.parametrize("small", [False, True])
def test_normal_equations_match_oracle(host, po, small):
    uv, p3, W, K, pose, _ = pnp_case(3, pn=11)
    if small:
        pose = pose.copy()
        pose[:3] = [3e-9, 1e-9, -4e-9]
    H, g, c = host.normal(pose, uv, p3, W, K)
    r, J = po.residuals_and_jacobian(pose, uv, p3, W, K)
    r, J = r.reshape(-1), J.reshape(-1, 6)
    JtJ = J.T @ J
    tri = np.array([JtJ[i, j] for i in range(6) for j in range(i, 6)])
    assert np.abs(H - tri).max() <= 1e-12 * np.abs(tri).max()
    assert np.abs(g - J.T @ r).max() <= 1e-12 * np.abs(J.T @ r).max()
    assert abs(c - 0.5 * r @ r) <= 1e-13 * c


def test_solver_follows_the_oracle_step_for_step(host, po):
    worst = 0.0
    for s in range(60):
        pn = int(np.random.default_rng(s).integers(5, 18))
        uv, p3, W, K, init, _ = pnp_case(200 + s, pn=pn, noise=2.0, pert=(0.3, 0.1) if s % 3 == 0 else (0.05, 0.02))
        want, info = po.uncertainty_pnp(uv, p3, W, K, init, return_info=True)
        got, (it, code) = host(uv, p3, W, K, init)
        assert (it, code) == (info["iterations"], info["termination"]), s
        worst = max(worst, np.abs(got - want).max())
    assert worst < 1e-12


def test_tight_tolerances_and_iteration_cap(host, po):
    uv, p3, W, K, init, _ = pnp_case(301, pn=9, noise=1.5, pert=(0.4, 0.15))
    want, info = po.uncertainty_pnp(uv, p3, W, K, init, max_num_iterations=200, function_tolerance=1e-16,
                                    gradient_tolerance=1e-14, parameter_tolerance=1e-16, return_info=True)
    got, (it, code) = host(uv, p3, W, K, init, mi=200, ft=1e-16, gt=1e-14, pt=1e-16)
    assert np.abs(got - want).max() < 1e-10
    got, (it, code) = host(uv, p3, W, K, init, mi=2)
    want, info = po.uncertainty_pnp(uv, p3, W, K, init, max_num_iterations=2, return_info=True)
    assert (it, code) == (info["iterations"], info["termination"]) and np.abs(got - want).max() < 1e-12


def test_degenerate_inputs(host, po):
    uv, p3, W, K, init, true_rt = pnp_case(302, noise=0.0)
    got, (it, code) = host(uv, p3, np.zeros_like(W), K, init)                  # all weights zero -> nothing to do
    assert it == 0 and code == po.CONVERGENCE_GRADIENT and np.array_equal(got, init)
    got, (it, code) = host(uv, p3, W, K, true_rt)                             # start at the optimum
    assert np.abs(got - true_rt).max() < 1e-9
    behind = init.copy()
    behind[5] = -0.05
    with np.errstate(all="ignore"):
        want, info = po.uncertainty_pnp(uv, p3, W, K, behind, return_info=True)
    got, (it, code) = host(uv, p3, W, K, behind)
    assert (it, code) == (info["iterations"], info["termination"])
    assert np.allclose(got, want, rtol=1e-9, atol=1e-9, equal_nan=True)
    Wn = W.copy()
    Wn[2, 0] = np.nan                                                          # NaN weight: both stop at once, x = init
    with np.errstate(all="ignore"):
        want, info = po.uncertainty_pnp(uv, p3, Wn, K, init, return_info=True)
    got, (it, code) = host(uv, p3, Wn, K, init)
    assert (it, code) == (info["iterations"], info["termination"]) and np.array_equal(got, init)
