#!/usr/bin/env python
"""Generates tests/golden/ceres_pnp.npz by running the reference's uncertainty-PnP through REAL Ceres 2.0 -- the
reference's own prebuilt lib/libceres.so.2.0.0 and its unmodified src/uncertainty_pnp.cpp, built into oracle/_ref/ceres/
by oracle/build_ceres_ref.py.  CPU only; needs /root/reference:

    python tests/golden/make_golden_ceres.py

Per problem: the inputs, Ceres' result, why and after how many iterations it stopped, the cost after every iteration, and a
`sensitivity`: how far the result moves when the same problem is started from init*(1+1e-13); `stable` = below 1e-10.  Long descents of
ill-conditioned problems (4-6 points, started far away) amplify a last-bit difference by ~10x per iteration
(profiles/r02_ceres_pin.md); no independent implementation can follow those, so bit-level checks use the stable ones.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import build_ceres_ref as ceres  # noqa: E402
from util import pnp_case  # noqa: E402

PN_MAX, TR_MAX = 24, 56


def problem_sets():
    sets = []
    for s in range(64):      # the batch of tests/test_gpu_pnp.py::test_batch_matches_*
        sets.append(("default", pnp_case(400 + s, pn=9, noise=2.0, pert=(0.3, 0.1) if s % 3 == 0 else (0.05, 0.02))[:5]))
    for s in range(60):      # the set of tests/test_pnp_host_core.py
        pn = int(np.random.default_rng(s).integers(5, 18))
        sets.append(("ragged", pnp_case(200 + s, pn=pn, noise=2.0, pert=(0.3, 0.1) if s % 3 == 0 else (0.05, 0.02))[:5]))
    for s in range(100):     # started far away: rejected steps, radius shrinking, iteration cap
        sets.append(("far", pnp_case(3000 + s, pn=int(4 + s % 20), noise=3.0, pert=(1.5, 0.5))[:5]))
    for s in range(40):      # initial pose behind the camera
        c = list(pnp_case(4000 + s, pn=9, noise=1.0)[:5])
        c[4] = c[4].copy()
        c[4][5] = -0.05
        sets.append(("behind", tuple(c)))
    for s in range(20):      # noise-free data started at the optimum
        c = pnp_case(5000 + s, pn=9, noise=0.0)
        sets.append(("optimum", (c[0], c[1], c[2], c[3], c[5])))
    return sets


def main():
    if ceres.build() is None:
        raise SystemExit("needs the reference checkout (/root/reference)")
    sets = problem_sets()
    n = len(sets)
    out = dict(kind=np.array([k for k, _ in sets]), pn=np.zeros(n, np.int32), pts2d=np.zeros((n, PN_MAX, 2)),
               pts3d=np.zeros((n, PN_MAX, 3)), wgt2d=np.zeros((n, PN_MAX, 3)), K=np.zeros((n, 3, 3)), init_rt=np.zeros((n, 6)),
               result_rt=np.zeros((n, 6)), entry_rt=np.zeros((n, 6)), reason=np.zeros(n, np.int32),
               termination_type=np.zeros(n, np.int32), iteration_summaries=np.zeros(n, np.int32),
               successful=np.zeros(n, np.int32), unsuccessful=np.zeros(n, np.int32), initial_cost=np.zeros(n),
               final_cost=np.zeros(n), cost_trace=np.full((n, TR_MAX), np.nan), step_ok=np.zeros((n, TR_MAX), np.uint8),
               stable=np.zeros(n, bool), sensitivity=np.zeros(n), linear_solver_type_used=np.zeros(n, np.int32))
    for i, (_, (uv, p3, W, K, init)) in enumerate(sets):
        pn = len(uv)
        with np.errstate(all="ignore"):
            res, info, tr = ceres.solve(uv, p3, W, K, init)
            ent = ceres.reference_entry(uv, p3, W, K, init)          # the reference's own C entry: same bits expected
            res2, _, _ = ceres.solve(uv, p3, W, K, init * (1.0 + 1e-13))
        assert np.array_equal(res, ent, equal_nan=True), i
        out["pn"][i] = pn
        out["pts2d"][i, :pn], out["pts3d"][i, :pn], out["wgt2d"][i, :pn], out["K"][i], out["init_rt"][i] = uv, p3, W, K, init
        out["result_rt"][i], out["entry_rt"][i] = res, ent
        out["reason"][i], out["termination_type"][i] = info["reason"], info["termination_type"]
        out["iteration_summaries"][i] = info["iteration_summaries"]
        out["successful"][i], out["unsuccessful"][i] = info["successful_steps"], info["unsuccessful_steps"]
        out["initial_cost"][i], out["final_cost"][i] = info["initial_cost"], info["final_cost"]
        out["linear_solver_type_used"][i] = info["linear_solver_type_used"]
        k = min(len(tr), TR_MAX)
        out["cost_trace"][i, :k] = tr[:k, 1]
        out["step_ok"][i, :k] = tr[:k, 7].astype(np.uint8)
        sens = np.abs(res - res2).max() if (np.isfinite(res).all() and np.isfinite(res2).all()) else np.inf
        out["sensitivity"][i] = sens
        out["stable"][i] = bool(sens < 1e-10)
    np.savez_compressed(os.path.join(HERE, "ceres_pnp.npz"), **out)
    for kind in ("default", "ragged", "far", "behind", "optimum"):
        m = out["kind"] == kind
        print(f"{kind:8s} n={m.sum():3d} stable={out['stable'][m].sum():3d} reasons={np.bincount(out['reason'][m], minlength=7).tolist()} "
              f"max unsuccessful={out['unsuccessful'][m].max()}")


if __name__ == "__main__":
    main()
