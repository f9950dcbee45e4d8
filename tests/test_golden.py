"""CPU: the oracle against golden vectors produced by the UNMODIFIED reference extension + Python
operator on a B200 (tests/golden/make_golden.py; regenerate with gpurun).  This is the oracle's pin:
hypotheses bit-equal, inlier bytes/counts equal, keypoints < 1e-3 px, covariance rtol 2e-3."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not generated yet")
    return np.load(path)


def test_kernels_against_reference_extension(oracle):
    g = _load("kernels.npz")
    direct, coords, idxs = g["direct"], g["coords"], g["idxs"]
    hyp = oracle.generate_hypothesis(direct, coords, idxs)
    assert np.array_equal(hyp.view(np.uint32), g["hyp"].view(np.uint32))
    assert np.array_equal(oracle.vote_count(direct, coords, g["hyp"], 0.99), g["counts_099"])
    assert np.array_equal(oracle.vote_count(direct, coords, g["hyp"], 0.999), g["counts_0999"])
    inl = np.zeros((48, 3, 600), dtype=np.uint8)
    oracle.voting_for_hypothesis(direct, coords, g["hyp"], inl, 0.99)
    assert np.array_equal(np.packbits(inl, axis=2), g["inliers_099_packed"])
    hyp_vp = oracle.generate_hypothesis(direct, coords, idxs, vanishing_point=True)
    assert np.array_equal(hyp_vp.view(np.uint32), g["hyp_vp"].view(np.uint32))
    inl = np.zeros((48, 3, 600), dtype=np.uint8)
    oracle.voting_for_hypothesis(direct, coords, g["hyp_vp"], inl, 0.999, vanishing_point=True)
    assert np.array_equal(np.packbits(inl, axis=2), g["inliers_vp_packed"])


@pytest.mark.parametrize("name", ["v3_plain.npz", "v3_thinned.npz"])
def test_v3_against_reference_operator(oracle, name):
    g = _load(name)
    sel = g["selection"] if g["selection"].size else None
    out = oracle.ransac_voting_layer_v3(g["mask"].astype(np.int64), g["vertex"], int(g["hn"]),
                                        inlier_thresh=float(g["thresh"]), max_num=int(g["max_num"]),
                                        idxs=g["idxs"], selection=sel)
    err = np.linalg.norm(out - g["kpt"], axis=-1).max()
    assert err < 1e-3, err


def test_distribution_against_reference_operator(oracle):
    g = _load("dist.npz")
    _, cov = oracle.estimate_voting_distribution_with_mean(g["mask"].astype(np.int64), g["vertex"], g["mean"],
                                                           round_hyp_num=int(g["round_hyp_num"]),
                                                           min_hyp_num=int(g["min_hyp_num"]), idxs=g["idxs"])
    assert np.allclose(cov, g["cov"], rtol=2e-3, atol=1e-3), np.abs(cov - g["cov"]).max()
