"""CPU: the uncertainty-PnP oracle (oracle/pnp_oracle.py) against independent checks.

The reference's minimiser (Ceres 2.0) cannot be built or loaded here, so the oracle's LM loop is a restatement of the
published algorithm (PARITY UNPINNED against Ceres, see the oracle's header).  What these tests pin:
  * the objective: residuals against an independent Rodrigues implementation, the forward-mode Jacobian against central
    differences (both branches of rotation.h);
  * the optimum: scipy.optimize.least_squares (MINPACK lmder) on the same residuals;
  * the convergence slack of Ceres' default tolerances (function_tolerance 1e-6): <= 2e-4 in pose units.
"""
import numpy as np
import pytest
from scipy.optimize import least_squares

from util import pnp_case


@pytest.fixture(scope="module")
def po():
    import pnp_oracle
    return pnp_oracle


def test_residuals_against_rotation_matrix(po):
    uv, p3, W, K, init, _ = pnp_case(1)
    R = po.rodrigues(init[:3])
    cam = p3 @ R.T + init[3:]
    d = np.stack([K[0, 0] * cam[:, 0] / cam[:, 2] + K[0, 2] - uv[:, 0], K[1, 1] * cam[:, 1] / cam[:, 2] + K[1, 2] - uv[:, 1]], 1)
    want = np.stack([W[:, 0] * d[:, 0] + W[:, 1] * d[:, 1], W[:, 1] * d[:, 0] + W[:, 2] * d[:, 1]], 1)
    assert np.abs(po.residuals(init, uv, p3, W, K) - want).max() < 1e-9


@pytest.mark.parametrize("small", [False, True])
def test_jacobian_against_central_differences(po, small):
    uv, p3, W, K, pose, _ = pnp_case(2)
    if small:
        pose = pose.copy()
        pose[:3] = [1e-9, -2e-9, 5e-10]          # theta^2 < epsilon: the first-order branch of rotation.h
    r, J = po.residuals_and_jacobian(pose, uv, p3, W, K)
    assert np.abs(r - po.residuals(pose, uv, p3, W, K)).max() < 1e-11
    num = np.empty_like(J)
    for k in range(6):
        e = np.zeros(6)
        e[k] = 1e-6
        num[:, :, k] = (po.residuals(pose + e, uv, p3, W, K) - po.residuals(pose - e, uv, p3, W, K)) / 2e-6
    assert np.abs(J - num).max() < 2e-8 * np.abs(J).max()


@pytest.mark.parametrize("seed", range(12))
def test_optimum_matches_scipy(po, seed):
    pn = int(np.random.default_rng(seed).integers(5, 18))
    uv, p3, W, K, init, _ = pnp_case(100 + seed, pn=pn, noise=2.0, pert=(0.3, 0.1) if seed % 3 == 0 else (0.05, 0.02))
    ref = least_squares(lambda p: po.residuals(p, uv, p3, W, K).ravel(), init, method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15)
    tight, info = po.uncertainty_pnp(uv, p3, W, K, init, max_num_iterations=200, function_tolerance=1e-16,
                                     gradient_tolerance=1e-14, parameter_tolerance=1e-16, return_info=True)
    assert np.abs(tight - ref.x).max() < 1e-8, info
    default, info = po.uncertainty_pnp(uv, p3, W, K, init, return_info=True)
    assert info["termination"] in (po.CONVERGENCE_FUNCTION, po.CONVERGENCE_PARAMETER, po.CONVERGENCE_GRADIENT)
    assert info["iterations"] <= 20
    assert np.abs(default - ref.x).max() < 2e-4           # what function_tolerance = 1e-6 leaves on the table


def test_termination_paths(po):
    uv, p3, W, K, init, true_rt = pnp_case(7, noise=0.0)
    x, info = po.uncertainty_pnp(uv, p3, W, K, true_rt, return_info=True)            # already optimal
    assert info["termination"] in (po.CONVERGENCE_GRADIENT, po.CONVERGENCE_FUNCTION, po.CONVERGENCE_PARAMETER)
    assert np.abs(x - true_rt).max() < 1e-9
    x, info = po.uncertainty_pnp(uv, p3, np.zeros_like(W), K, init, return_info=True)   # all weights zero (:121-122 of pvnet.py)
    assert info["termination"] == po.CONVERGENCE_GRADIENT and info["iterations"] == 0 and np.array_equal(x, init)
    x, info = po.uncertainty_pnp(uv, p3, W, K, init, max_num_iterations=1, return_info=True)
    assert info["iterations"] == 1
    behind = init.copy()
    behind[5] = -0.05                                                                  # points behind / at the camera
    with np.errstate(all="ignore"):
        x, info = po.uncertainty_pnp(uv, p3, W, K, behind, return_info=True)
    assert info["termination"] in range(1, 7) and x.shape == (6,)
