"""GPU: the CUDA path (through the C ABI) against the golden vectors of the reference extension."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not generated yet")
    return np.load(path)


def _t(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def test_kernels_against_golden(pvb):
    g = _load("kernels.npz")
    d, c, i = _t(g["direct"]), _t(g["coords"]), _t(g["idxs"])
    hyp = pvb.ransac_voting.generate_hypothesis(d, c, i)
    assert np.array_equal(hyp.cpu().numpy().view(np.uint32), g["hyp"].view(np.uint32))
    assert np.array_equal(pvb.ransac_voting.vote_count(d, c, hyp, 0.99).cpu().numpy(), g["counts_099"])
    assert np.array_equal(pvb.ransac_voting.vote_count(d, c, hyp, 0.999).cpu().numpy(), g["counts_0999"])
    inl = torch.zeros((48, 3, 600), dtype=torch.uint8, device="cuda")
    pvb.ransac_voting.voting_for_hypothesis(d, c, hyp, inl, 0.99)
    assert np.array_equal(np.packbits(inl.cpu().numpy(), axis=2), g["inliers_099_packed"])
    hyp_vp = pvb.ransac_voting.generate_hypothesis_vanishing_point(d, c, i)
    assert np.array_equal(hyp_vp.cpu().numpy().view(np.uint32), g["hyp_vp"].view(np.uint32))
    inl.zero_()
    pvb.ransac_voting.voting_for_hypothesis_vanishing_point(d, c, hyp_vp, inl, 0.999)
    assert np.array_equal(np.packbits(inl.cpu().numpy(), axis=2), g["inliers_vp_packed"])


@pytest.mark.parametrize("name", ["v3_plain.npz", "v3_thinned.npz"])
def test_v3_against_golden(pvb, name):
    g = _load(name)
    sel = _t(g["selection"]) if g["selection"].size else None
    out = pvb.ransac_voting_layer_v3(_t(g["mask"]), _t(g["vertex"]), int(g["hn"]), inlier_thresh=float(g["thresh"]),
                                     max_num=int(g["max_num"]), idxs=_t(g["idxs"]), selection=sel)
    err = np.linalg.norm(out.cpu().numpy() - g["kpt"], axis=-1).max()
    assert err < 1e-3, err


def test_distribution_against_golden(pvb):
    g = _load("dist.npz")
    _, cov = pvb.estimate_voting_distribution_with_mean(_t(g["mask"]), _t(g["vertex"]), _t(g["mean"]),
                                                        round_hyp_num=int(g["round_hyp_num"]),
                                                        min_hyp_num=int(g["min_hyp_num"]), idxs=_t(g["idxs"]))
    assert np.allclose(cov.cpu().numpy(), g["cov"], rtol=2e-3, atol=1e-3)
