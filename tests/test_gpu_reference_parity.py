"""GPU parity against the reference itself: the unmodified clean-pvnet CUDA extension compiled for
sm_100 by oracle/build_ref.py, run in the same process on the same tensors.

  kernel level : generate_hypothesis bit-equal, voting_for_hypothesis bytes equal, counts equal
  op level     : ransac_voting_layer_v3 / estimate_voting_distribution_with_mean under the same
                 torch.manual_seed (rng="torch" replays the reference's generator consumption):
                 keypoint L2 error < 1e-3 px (north-star bar), covariance rtol 2e-3.
"""
import numpy as np
import pytest
import torch

from refload import load_reference
from util import bits_equal, cuda, field_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    try:
        return load_reference()
    except Exception as e:   # pragma: no cover
        pytest.skip(f"reference extension unavailable: {e}")


def _inputs(cfg, **kw):
    from clean_pvnet_b200 import synth
    return synth.make_inputs(cfg, device="cuda", **kw)


@pytest.mark.parametrize("tn,vn,hn,seed", [(3000, 3, 128, 0), (4096, 9, 512, 1), (700, 1, 64, 2)])
def test_kernels_match_reference_extension(pvb, ref, tn, vn, hn, seed):
    ext, _ = ref
    direct, coords, idxs, _ = field_case(tn, vn, hn, seed)
    d, c, i = cuda(direct, coords, idxs)
    want_h = ext.generate_hypothesis(d, c, i)
    got_h = pvb.ransac_voting.generate_hypothesis(d, c, i)
    assert bits_equal(got_h.cpu().numpy(), want_h.cpu().numpy())
    for thresh in (0.99, 0.999):
        want_i = torch.zeros((hn, vn, tn), dtype=torch.uint8, device="cuda")
        ext.voting_for_hypothesis(d, c, want_h, want_i, thresh)
        got_i = torch.zeros_like(want_i)
        pvb.ransac_voting.voting_for_hypothesis(d, c, want_h, got_i, thresh)
        assert torch.equal(got_i, want_i)
        counts = pvb.ransac_voting.vote_count(d, c, want_h, thresh)
        assert torch.equal(counts, want_i.sum(dim=2, dtype=torch.int32))


def test_vanishing_point_kernels_match_reference_extension(pvb, ref):
    ext, _ = ref
    direct, coords, idxs, _ = field_case(2000, 3, 128, 5)
    d, c, i = cuda(direct, coords, idxs)
    want_h = ext.generate_hypothesis_vanishing_point(d, c, i)
    got_h = pvb.ransac_voting.generate_hypothesis_vanishing_point(d, c, i)
    assert bits_equal(got_h.cpu().numpy(), want_h.cpu().numpy())
    want_i = torch.zeros((128, 3, 2000), dtype=torch.uint8, device="cuda")
    ext.voting_for_hypothesis_vanishing_point(d, c, want_h, want_i, 0.999)
    got_i = torch.zeros_like(want_i)
    pvb.ransac_voting.voting_for_hypothesis_vanishing_point(d, c, want_h, got_i, 0.999)
    assert torch.equal(got_i, want_i)


@pytest.mark.parametrize("cfg,hn,max_num,seed", [("small", 64, 30000, 0), ("small", 128, 700, 1), ("tiny", 32, 30000, 2)])
def test_v3_matches_reference_under_same_seed(pvb, ref, cfg, hn, max_num, seed):
    _, gpu = ref
    mask, vertex, _ = _inputs(cfg, seed=100 + seed)
    torch.manual_seed(seed)
    want = gpu.ransac_voting_layer_v3(mask, vertex, hn, inlier_thresh=0.99, max_num=max_num)
    torch.manual_seed(seed)
    got = pvb.ransac_voting_layer_v3(mask, vertex, hn, inlier_thresh=0.99, max_num=max_num, rng="torch")
    err = (got - want).norm(dim=-1).max().item()
    assert err < 1e-3, err


def _exact_refit(ext, dbg, thresh):
    """The reference's refit formula (ransac_voting_gpu.py:177-196) in float64 on the inlier set the
    reference extension itself produces for the winning hypotheses -- the value both implementations
    approximate (the reference in fp32 through cuBLAS matmul + torch.sum + LU)."""
    B, K = dbg["win"].shape[:2]
    out = torch.zeros((B, K, 2), dtype=torch.float64, device="cuda")
    for b in range(B):
        tn = int(dbg["tn"][b])
        direct = dbg["dirs"][b, :, :tn].permute(1, 0, 2).contiguous()
        coords = dbg["xy"][b, :tn].contiguous()
        inl = torch.zeros((1, K, tn), dtype=torch.uint8, device="cuda")
        ext.voting_for_hypothesis(direct, coords, dbg["win"][b][None].contiguous(), inl, thresh)
        w = inl[0].double()                                           # [K,tn]
        normal = torch.stack([direct[:, :, 1], -direct[:, :, 0]], dim=-1).double().permute(1, 0, 2) * w[:, :, None]
        bb = (normal * coords.double()[None]).sum(2)                  # [K,tn]
        ATA = normal.transpose(1, 2) @ normal
        ATb = (normal * bb[:, :, None]).sum(1)
        out[b] = torch.linalg.solve(ATA, ATb[:, :, None])[:, :, 0]
    return out


def _reference_refit_lines(ext, gpu, direct, coords, win, thresh):
    """ransac_voting_gpu.py:177-196 as the reference executes them (fp32 torch ops: matmul, sum, its own b_inv) on a given
    pixel ORDER -- the winner's inlier set does not depend on the order, the fp32 sums do."""
    tn, vn = direct.shape[0], direct.shape[1]
    normal = torch.zeros_like(direct)
    normal[:, :, 0] = direct[:, :, 1]
    normal[:, :, 1] = -direct[:, :, 0]
    inl = torch.zeros([1, vn, tn], dtype=torch.uint8, device=direct.device)
    ext.voting_for_hypothesis(direct, coords, win[None].contiguous(), inl, thresh)
    inl = torch.squeeze(inl.float(), 0)
    normal = normal.permute(1, 0, 2) * torch.unsqueeze(inl, 2)
    b = torch.sum(normal * torch.unsqueeze(coords, 0), 2)
    ATA = torch.matmul(normal.permute(0, 2, 1), normal)
    ATb = torch.sum(normal * torch.unsqueeze(b, 2), 1)
    return torch.matmul(gpu.b_inv(ATA), torch.unsqueeze(ATb, 2))[:, :, 0], inl.sum(1)


@pytest.mark.parametrize("layout", ["planar", "interleaved"])
def test_v3_full_size_matches_reference(pvb, ref, layout):
    """The FULL cfg-2 batch (B=16, 480x640, K=9, hn=512, thinning: fg ~ 92k > 30000), both vertex layouts, same seed.

    With ~22 000 inliers per keypoint the reference's fp32 normal equations (cuBLAS matmul + torch.sum,
    ransac_voting_gpu.py:189-193) carry their own rounding noise, largest on the ill-conditioned out-of-image keypoint;
    `<1e-3 px vs reference` is therefore only defined up to the reference's distance to itself.  That distance is measured
    here: the reference's own refit lines re-run on a PERMUTED pixel order (same inlier set, same formula, same fp32 ops).
    Pinned: identical inlier sets; ours within 1e-4 px of the exact (float64) value of the reference's formula; the
    ours-vs-reference gap explained by the reference's distance to that value; and ours-vs-reference no larger than a small
    multiple of reference-vs-itself.  The numbers are written to gpurun_out/ for DESIGN.md / BASELINE.md."""
    ext, gpu = ref
    mask, vertex, _ = _inputs("cfg2", seed=77, layout=layout)
    torch.manual_seed(3)
    want = gpu.ransac_voting_layer_v3(mask, vertex, 512, inlier_thresh=0.99)
    torch.manual_seed(3)
    got, dbg = pvb.ransac_voting_layer_v3(mask, vertex, 512, inlier_thresh=0.99, rng="torch", debug=True)
    exact = _exact_refit(ext, dbg, 0.99)
    ours_vs_exact = (got.double() - exact).norm(dim=-1)
    ref_vs_exact = (want.double() - exact).norm(dim=-1)
    ours_vs_ref = (got - want).norm(dim=-1)
    # the reference against itself: identity order must reproduce its output, a permuted order shows its fp32 spread
    B, K = got.shape[:2]
    same_order = torch.zeros_like(want)
    permuted = torch.zeros_like(want)
    g = torch.Generator(device="cuda").manual_seed(5)
    for b in range(B):
        tn = int(dbg["tn"][b])
        direct = dbg["dirs"][b, :, :tn].permute(1, 0, 2).contiguous()
        coords = dbg["xy"][b, :tn].contiguous()
        same_order[b], n0 = _reference_refit_lines(ext, gpu, direct, coords, dbg["win"][b], 0.99)
        perm = torch.randperm(tn, generator=g, device="cuda")
        permuted[b], n1 = _reference_refit_lines(ext, gpu, direct[perm].contiguous(), coords[perm].contiguous(), dbg["win"][b], 0.99)
        assert torch.equal(n0, n1)                                   # the inlier SET is order independent
    ref_vs_itself = (permuted - want).norm(dim=-1)
    assert (same_order - want).norm(dim=-1).max().item() < 1e-4     # the helper is the reference's own computation
    assert ours_vs_exact.max().item() < 1e-4, ours_vs_exact.max().item()
    assert (ours_vs_ref.double() <= ref_vs_exact + 2e-4).all(), (ours_vs_ref, ref_vs_exact)
    spread = max(ref_vs_itself.max().item(), ref_vs_exact.max().item())
    assert ours_vs_ref.max().item() <= max(1e-3, 3.0 * spread), (ours_vs_ref.max().item(), spread)
    inside = ours_vs_ref[:, :-1]                                     # keypoints inside the image (well conditioned)
    report = dict(layout=layout, images=B, ours_vs_ref_max=ours_vs_ref.max().item(), ours_vs_ref_inside_max=inside.max().item(),
                  ref_vs_itself_permuted_max=ref_vs_itself.max().item(), ref_vs_itself_inside_max=ref_vs_itself[:, :-1].max().item(),
                  ref_vs_exact_max=ref_vs_exact.max().item(), ours_vs_exact_max=ours_vs_exact.max().item(),
                  ours_below_1e3_fraction=float((ours_vs_ref < 1e-3).float().mean().item()))
    import json
    import os
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, f"r02_parity_cfg2_{layout}.json"), "w") as fh:
            json.dump(report, fh, indent=1)


def test_distribution_matches_reference_under_same_seed(pvb, ref):
    _, gpu = ref
    mask, vertex, _ = _inputs("small", seed=200)
    mean = pvb.ransac_voting_layer_v3(mask, vertex, 64, inlier_thresh=0.99, seed=1)
    torch.manual_seed(9)
    _, want = gpu.estimate_voting_distribution_with_mean(mask, vertex, mean.clone(), round_hyp_num=64, min_hyp_num=512)
    torch.manual_seed(9)
    _, got = pvb.estimate_voting_distribution_with_mean(mask, vertex, mean, round_hyp_num=64, min_hyp_num=512,
                                                        rng="torch")
    assert torch.allclose(got, want, rtol=2e-3, atol=1e-3), (got - want).abs().max().item()


def test_distribution_with_thinning_matches_reference(pvb, ref):
    """fg > max_num: the reference thins, recomputes `foreground` (:219-223) and divides counts by it."""
    _, gpu = ref
    mask, vertex, _ = _inputs("small", seed=201)
    mean = pvb.ransac_voting_layer_v3(mask, vertex, 64, inlier_thresh=0.99, seed=1)
    torch.manual_seed(10)
    _, want = gpu.estimate_voting_distribution_with_mean(mask, vertex, mean.clone(), round_hyp_num=64, min_hyp_num=256,
                                                         max_num=700)
    torch.manual_seed(10)
    _, got = pvb.estimate_voting_distribution_with_mean(mask, vertex, mean, round_hyp_num=64, min_hyp_num=256,
                                                        max_num=700, rng="torch")
    assert torch.allclose(got, want, rtol=2e-3, atol=1e-3), (got - want).abs().max().item()


def test_v1_layer_matches_reference(pvb, ref):
    """ransac_voting_layer (v1, torch.inverse instead of b_inv) -- imported by resnet18.py:5."""
    _, gpu = ref
    mask, vertex, _ = _inputs("small", seed=202)
    torch.manual_seed(4)
    want = gpu.ransac_voting_layer(mask, vertex, 64, inlier_thresh=0.99)
    torch.manual_seed(4)
    got = pvb.ransac_voting_layer(mask, vertex, 64, inlier_thresh=0.99, rng="torch")
    assert (got - want).norm(dim=-1).max().item() < 1e-3


def test_production_call_pattern(pvb, ref):
    """decode_keypoint's two call patterns (resnet18.py:70-76) on the strided NCHW view with an argmax mask."""
    _, gpu = ref
    from clean_pvnet_b200 import synth
    mask, vertex, _ = synth.make_inputs("small", device="cuda", seed=203, layout="planar")
    seg = torch.stack([1.0 - mask.float(), mask.float()], dim=1)           # [B,2,H,W] logits
    amask = torch.argmax(seg, 1)                                           # int64, as resnet18.py:69
    torch.manual_seed(6)
    want = gpu.ransac_voting_layer_v3(amask, vertex, 128, inlier_thresh=0.99, max_num=100)
    torch.manual_seed(6)
    got = pvb.ransac_voting_layer_v3(amask, vertex, 128, inlier_thresh=0.99, max_num=100, rng="torch")
    assert (got - want).norm(dim=-1).max().item() < 1e-3
    torch.manual_seed(7)
    mean_w = gpu.ransac_voting_layer_v3(amask, vertex, 512, inlier_thresh=0.99)
    _, var_w = gpu.estimate_voting_distribution_with_mean(amask, vertex, mean_w)
    torch.manual_seed(7)
    mean_g = pvb.ransac_voting_layer_v3(amask, vertex, 512, inlier_thresh=0.99, rng="torch")
    _, var_g = pvb.estimate_voting_distribution_with_mean(amask, vertex, mean_g, rng="torch")
    assert (mean_g - mean_w).norm(dim=-1).max().item() < 1e-3
    assert torch.allclose(var_g, var_w, rtol=5e-3, atol=2e-3), (var_g - var_w).abs().max().item()


def test_philox_mode_is_statistically_equivalent(pvb, ref):
    """Default (philox) sampling is a different random stream, not a different estimator."""
    _, gpu = ref
    mask, vertex, kp = _inputs("small", seed=300)
    torch.manual_seed(0)
    want = gpu.ransac_voting_layer_v3(mask, vertex, 128, inlier_thresh=0.99)
    got = pvb.ransac_voting_layer_v3(mask, vertex, 128, inlier_thresh=0.99, seed=0)
    # both land on the same inlier consensus: sub-pixel agreement on in-image keypoints
    assert (got - want)[:, :-1].norm(dim=-1).max().item() < 0.75
