"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, float64) of the reference's uncertainty-PnP refinement.

Reference being restated
    lib/csrc/uncertainty_pnp/src/uncertainty_pnp.cpp:7-55   the residual functor (angle-axis rotate, translate, pinhole
                                                            projection, 2x2 symmetric weight)
    lib/csrc/uncertainty_pnp/src/uncertainty_pnp.cpp:61-92  `uncertainty_pnp(...)`: one ceres::Problem, 6 parameters, one
                                                            AutoDiffCostFunction<.,2,6> per point, default Solver::Options
                                                            except linear_solver_type = DENSE_SCHUR, ceres::Solve.
    lib/csrc/uncertainty_pnp/un_pnp_utils.py:6-57           the Python caller (P3P initialisation with OpenCV, then the C entry)

The minimiser is a third-party dependency that is NOT in /root/reference as source: Ceres Solver 2.0.0, vendored as headers
under lib/csrc/uncertainty_pnp/include/ceres plus the prebuilt lib/libceres.so.2.0.0.  PARITY PINNED AGAINST THAT BINARY:
oracle/build_ceres_ref.py loads it here (its six missing back-end libraries -- SuiteSparse, CXSparse, LAPACK/BLAS, gflags,
libunwind -- are closed with abort-stubs; Ceres never calls them with DENSE_SCHUR + Eigen) and builds the reference's
unmodified src/uncertainty_pnp.cpp against it; tests/golden/make_golden_ceres.py recorded 284 problems it solved
(tests/golden/ceres_pnp.npz) and tests/test_ceres_golden.py holds this restatement to it: same stop reason, same number of
iterations, same cost after every iteration, same pose to 1e-9 on every problem an independent implementation can follow
(236 of them, incl. descents with up to 20 rejected steps and the 50-iteration cap; the rest are ill-conditioned 4-6 point
problems whose trajectories amplify a last-bit difference ~10x per iteration, measured by re-running Ceres itself from
init*(1+1e-13)).
What is restated is Ceres' TRUST_REGION / LEVENBERG_MARQUARDT loop with the default options (include/ceres/solver.h of the
vendored headers):
    max_num_iterations 50, function_tolerance 1e-6, gradient_tolerance 1e-10, parameter_tolerance 1e-8,
    initial_trust_region_radius 1e4, max 1e16, min 1e-32, min_relative_decrease 1e-3, min_lm_diagonal 1e-6,
    max_lm_diagonal 1e32, jacobi_scaling on, monotonic steps, max_num_consecutive_invalid_steps 5.
The constants of the loop can be read in source form in the vendored header-only `include/ceres/tiny_solver.h:171-290`, the same
authors' compact LM (Jacobi scaling 1/(1+|col|), LM diagonal sqrt(clamp(JtJ_ii, 1e-6, 1e32)/radius), radius /= max(1/3, 1-(2rho-1)^3)
on acceptance, radius /= v, v *= 2 on rejection); the full minimiser adds min_relative_decrease, the function-tolerance test
(which leaves the last candidate UNAPPLIED -- confirmed against the binary), invalid-step handling and evaluates the gradient
test on the unscaled gradient.  With one 6-parameter block DENSE_SCHUR eliminates it as the only e-block, i.e. solves the
6x6 regularised normal equations by Cholesky -- what `uncertainty_pnp` below does.
Also pinned independently: the objective (against ceres/rotation.h:563-607 semantics and a finite-difference check) and the
optimum (tests/test_pnp_oracle.py compares with scipy.optimize.least_squares on the same residuals).
"""
import numpy as np

EPS = np.finfo(np.float64).eps


def rotate_point(aa, p):
    """ceres::AngleAxisRotatePoint (include/ceres/rotation.h:563-607): Rodrigues away from zero, first-order near zero."""
    theta2 = float(aa @ aa)
    if theta2 > EPS:
        theta = np.sqrt(theta2)
        c, s = np.cos(theta), np.sin(theta)
        w = aa / theta
        return p * c + np.cross(w, p) * s + w * (w @ p) * (1.0 - c)
    return p + np.cross(aa, p)


def residuals(pose, pts2d, pts3d, wgt2d, K):
    """uncertainty_pnp.cpp:19-37 for every point; returns [pn,2]."""
    fx, fy, px, py = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    out = np.empty((pts2d.shape[0], 2))
    for i in range(pts2d.shape[0]):
        t = rotate_point(pose[:3], pts3d[i]) + pose[3:]
        with np.errstate(all="ignore"):
            dx = fx * t[0] / t[2] + px - pts2d[i, 0]
            dy = fy * t[1] / t[2] + py - pts2d[i, 1]
        out[i, 0] = wgt2d[i, 0] * dx + wgt2d[i, 1] * dy
        out[i, 1] = wgt2d[i, 1] * dx + wgt2d[i, 2] * dy
    return out


def _dual_rotate(aa, daa, p):
    """value and 3x6 Jacobian of rotate_point w.r.t. the 6 pose parameters (forward mode, like the AutoDiff functor;
    `daa` is the 3x6 seed of the angle-axis part).  Same branch structure as rotation.h, so the derivative at the
    small-angle branch is the one Ceres' jets produce."""
    theta2 = float(aa @ aa)
    if theta2 > EPS:
        dtheta2 = 2.0 * (aa @ daa)                       # [6]
        theta = np.sqrt(theta2)
        dtheta = dtheta2 / (2.0 * theta)
        c, s = np.cos(theta), np.sin(theta)
        dc, ds = -s * dtheta, c * dtheta
        ti = 1.0 / theta
        dti = -dtheta / theta2
        w = aa * ti
        dw = daa * ti + np.outer(aa, dti)                # [3,6]
        wxp = np.cross(w, p)
        dwxp = np.stack([dw[1] * p[2] - dw[2] * p[1], dw[2] * p[0] - dw[0] * p[2], dw[0] * p[1] - dw[1] * p[0]])
        wp = float(w @ p)
        dwp = p @ dw
        tmp = wp * (1.0 - c)
        dtmp = dwp * (1.0 - c) - wp * dc
        val = p * c + wxp * s + w * tmp
        jac = np.outer(p, dc) + dwxp * s + np.outer(wxp, ds) + dw * tmp + np.outer(w, dtmp)
        return val, jac
    val = p + np.cross(aa, p)
    jac = np.stack([daa[1] * p[2] - daa[2] * p[1], daa[2] * p[0] - daa[0] * p[2], daa[0] * p[1] - daa[1] * p[0]])
    return val, jac


def residuals_and_jacobian(pose, pts2d, pts3d, wgt2d, K):
    """r [pn,2] and J [pn,2,6] = d r / d pose, the quantities Ceres' AutoDiffCostFunction<.,2,6> hands to the solver."""
    fx, fy, px, py = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    pn = pts2d.shape[0]
    r = np.empty((pn, 2))
    J = np.empty((pn, 2, 6))
    seed = np.zeros((3, 6))
    seed[0, 0] = seed[1, 1] = seed[2, 2] = 1.0
    for i in range(pn):
        t, dt = _dual_rotate(pose[:3], seed, pts3d[i])
        t = t + pose[3:]
        dt = dt.copy()
        dt[0, 3] += 1.0; dt[1, 4] += 1.0; dt[2, 5] += 1.0
        with np.errstate(all="ignore"):
            iz = 1.0 / t[2]
            u, v = fx * t[0] * iz, fy * t[1] * iz
            du = fx * (dt[0] * iz - t[0] * iz * iz * dt[2])
            dv = fy * (dt[1] * iz - t[1] * iz * iz * dt[2])
            dx, dy = u + px - pts2d[i, 0], v + py - pts2d[i, 1]
        r[i, 0] = wgt2d[i, 0] * dx + wgt2d[i, 1] * dy
        r[i, 1] = wgt2d[i, 1] * dx + wgt2d[i, 2] * dy
        J[i, 0] = wgt2d[i, 0] * du + wgt2d[i, 1] * dv
        J[i, 1] = wgt2d[i, 1] * du + wgt2d[i, 2] * dv
    return r, J


# termination codes (shared with the CUDA implementation's `info` output)
CONVERGENCE_GRADIENT, CONVERGENCE_PARAMETER, CONVERGENCE_FUNCTION, CONVERGENCE_RADIUS, NO_CONVERGENCE, FAILURE = 1, 2, 3, 4, 5, 6


def uncertainty_pnp(pts2d, pts3d, wgt2d, K, init_rt, max_num_iterations=50, function_tolerance=1e-6,
                    gradient_tolerance=1e-10, parameter_tolerance=1e-8, return_info=False):
    """The C entry `uncertainty_pnp(pts2d, pts3d, wgt2d, K, init_rt, result_rt, pn)` (uncertainty_pnp.cpp:61-92):
    Levenberg-Marquardt trust-region minimisation of 0.5*sum r^2 from init_rt; returns result_rt [6]."""
    pts2d = np.asarray(pts2d, np.float64); pts3d = np.asarray(pts3d, np.float64)
    wgt2d = np.asarray(wgt2d, np.float64); K = np.asarray(K, np.float64).reshape(3, 3)
    x = np.asarray(init_rt, np.float64).reshape(6).copy()

    def evaluate(p):
        r, J = residuals_and_jacobian(p, pts2d, pts3d, wgt2d, K)
        r = r.reshape(-1); J = J.reshape(-1, 6)
        with np.errstate(all="ignore"):
            return 0.5 * float(r @ r), J.T @ J, J.T @ r

    cost, H, g = evaluate(x)
    x_norm = np.linalg.norm(x)
    scale = 1.0 / (1.0 + np.sqrt(np.diag(H)))            # jacobi scaling, fixed at iteration 0
    grad_max = np.max(np.abs(g))
    radius, decrease_factor, reuse_diagonal = 1e4, 2.0, False
    diag = None
    it, invalid = 0, 0
    code = None
    while True:
        # FinalizeIterationAndCheckIfMinimizerCanContinue
        if it >= max_num_iterations:
            code = NO_CONVERGENCE; break
        if not (grad_max > gradient_tolerance):
            code = CONVERGENCE_GRADIENT; break
        if radius < 1e-32:
            code = CONVERGENCE_RADIUS; break
        it += 1
        # LevenbergMarquardtStrategy::ComputeStep on the column-scaled Jacobian
        Hs = H * np.outer(scale, scale)
        bs = g * scale
        if not reuse_diagonal:
            diag = np.clip(np.diag(Hs), 1e-6, 1e32)
        A = Hs + np.diag(diag / radius)
        ok = True
        try:
            L = np.linalg.cholesky(A)
            y = np.linalg.solve(L.T, np.linalg.solve(L, bs))
            step = -y
            ok = bool(np.all(np.isfinite(step)))
        except np.linalg.LinAlgError:
            ok = False
        model_cost_change = -(step @ bs) - 0.5 * (step @ Hs @ step) if ok else -1.0
        if not (model_cost_change > 0.0):
            invalid += 1
            if invalid >= 5:
                code = FAILURE; break
            radius *= 0.5; reuse_diagonal = True          # StepIsInvalid
            continue
        invalid = 0
        delta = step * scale
        cand = x + delta
        ccost, cH, cg = evaluate(cand)
        if np.linalg.norm(delta) <= parameter_tolerance * (x_norm + parameter_tolerance):
            code = CONVERGENCE_PARAMETER; break           # the candidate is not adopted
        cost_change = cost - ccost
        if abs(cost_change) <= function_tolerance * cost:
            code = CONVERGENCE_FUNCTION; break            # the candidate is not adopted
        rho = cost_change / model_cost_change
        if rho > 1e-3:                                    # StepAccepted
            x, cost, H, g = cand, ccost, cH, cg
            x_norm = np.linalg.norm(x)
            grad_max = np.max(np.abs(g))
            radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3))
            decrease_factor, reuse_diagonal = 2.0, False
        else:                                             # StepRejected (also for a non-finite candidate cost)
            radius /= decrease_factor
            decrease_factor *= 2.0
            reuse_diagonal = True
    if return_info:
        return x, {"iterations": it, "termination": code, "cost": cost}
    return x


def rodrigues(aa):
    """3x3 rotation matrix of an angle-axis vector (what cv2.Rodrigues returns, un_pnp_utils.py:55)."""
    R = np.empty((3, 3))
    for j in range(3):
        e = np.zeros(3); e[j] = 1.0
        theta2 = float(aa @ aa)
        if theta2 > 0.0:
            theta = np.sqrt(theta2)
            w = aa / theta
            R[:, j] = e * np.cos(theta) + np.cross(w, e) * np.sin(theta) + w * (w @ e) * (1.0 - np.cos(theta))
        else:
            R[:, j] = e
    return R
