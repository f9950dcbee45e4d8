"""Twins of the reference pybind module `ransac_voting`
(lib/csrc/ransac_voting/src/ransac_voting.cpp:102-107), same names, argument order and error
behaviour (non-CUDA / non-contiguous tensors raise RuntimeError like CHECK_INPUT, .cpp:7-9),
backed by the sm_100a kernels in csrc/compat.cu through the C ABI.

    generate_hypothesis(direct[tn,vn,2] f32, coords[tn,2] f32, idxs[hn,vn,2] i32) -> hyp[hn,vn,2] f32
    voting_for_hypothesis(direct, coords, hypo_pts[hn,vn,2], inliers[hn,vn,tn] u8 (in/out), thresh) -> None
    generate_hypothesis_vanishing_point(...) -> hyp[hn,vn,3]
    voting_for_hypothesis_vanishing_point(direct, coords, hypo_pts[hn,vn,3], inliers, thresh) -> None

Extra (not in the reference): vote_count(direct, coords, hypo_pts, thresh) -> int32[hn,vn], the
fused voting_for_hypothesis + torch.sum(dim=2) (ransac_voting_gpu.py:156-159) run by the layer's
own vote kernel.
"""
import torch

from . import _lib


def _check_input(x, name, dtype):
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not x.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if x.dtype != dtype:
        raise RuntimeError(f"{name} must have dtype {dtype}, got {x.dtype}")


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def _shapes(direct, coords, third, last):
    if direct.dim() != 3 or direct.size(2) != 2:
        raise RuntimeError("direct must be [tn,vn,2]")
    tn, vn = direct.size(0), direct.size(1)
    if coords.dim() != 2 or coords.size(0) != tn or coords.size(1) != 2:
        raise RuntimeError("coords must be [tn,2]")
    if third.dim() != 3 or third.size(1) != vn or third.size(2) != last:
        raise RuntimeError(f"expected [hn,vn,{last}] tensor")
    return tn, vn, third.size(0)


def _generate(direct, coords, idxs, vanishing):
    _check_input(direct, "direct", torch.float32)
    _check_input(coords, "coords", torch.float32)
    _check_input(idxs, "idxs", torch.int32)
    tn, vn, hn = _shapes(direct, coords, idxs, 2)
    lib = _lib.load()
    with torch.cuda.device(direct.device):
        hyp = torch.empty((hn, vn, 3 if vanishing else 2), dtype=torch.float32, device=direct.device)
        fn = lib.pvb_generate_hypothesis_vanishing_point if vanishing else lib.pvb_generate_hypothesis
        _lib.check(fn(direct.data_ptr(), coords.data_ptr(), idxs.data_ptr(), hyp.data_ptr(), tn, vn, hn,
                      _stream(direct.device)))
    return hyp


def _vote(direct, coords, hypo_pts, inliers, inlier_thresh, vanishing):
    _check_input(direct, "direct", torch.float32)
    _check_input(coords, "coords", torch.float32)
    _check_input(hypo_pts, "hypo_pts", torch.float32)
    _check_input(inliers, "inliers", torch.uint8)
    tn, vn, hn = _shapes(direct, coords, hypo_pts, 3 if vanishing else 2)
    if tuple(inliers.shape) != (hn, vn, tn):
        raise RuntimeError("inliers must be [hn,vn,tn]")
    lib = _lib.load()
    with torch.cuda.device(direct.device):
        fn = lib.pvb_voting_for_hypothesis_vanishing_point if vanishing else lib.pvb_voting_for_hypothesis
        _lib.check(fn(direct.data_ptr(), coords.data_ptr(), hypo_pts.data_ptr(), inliers.data_ptr(), tn, vn, hn,
                      float(inlier_thresh), _stream(direct.device)))


def generate_hypothesis(direct, coords, idxs):
    """ransac_voting.cpp:20-31 / ransac_voting_kernel.cu:11-86."""
    return _generate(direct, coords, idxs, False)


def voting_for_hypothesis(direct, coords, hypo_pts, inliers, inlier_thresh):
    """ransac_voting.cpp:41-55 / ransac_voting_kernel.cu:88-167.  `inliers` is updated in place."""
    _vote(direct, coords, hypo_pts, inliers, inlier_thresh, False)


def generate_hypothesis_vanishing_point(direct, coords, idxs):
    """ransac_voting.cpp:64-75 / ransac_voting_kernel.cu:170-266."""
    return _generate(direct, coords, idxs, True)


def voting_for_hypothesis_vanishing_point(direct, coords, hypo_pts, inliers, inlier_thresh):
    """ransac_voting.cpp:85-99 / ransac_voting_kernel.cu:268-351."""
    _vote(direct, coords, hypo_pts, inliers, inlier_thresh, True)


def vote_count(direct, coords, hypo_pts, inlier_thresh):
    """Inlier counts int32 [hn,vn] == voting_for_hypothesis(...) followed by sum over tn."""
    _check_input(direct, "direct", torch.float32)
    _check_input(coords, "coords", torch.float32)
    _check_input(hypo_pts, "hypo_pts", torch.float32)
    tn, vn, hn = _shapes(direct, coords, hypo_pts, 2)
    lib = _lib.load()
    with torch.cuda.device(direct.device):
        counts = torch.empty((hn, vn), dtype=torch.int32, device=direct.device)
        nbytes = lib.pvb_vote_count_workspace_bytes(tn, vn, hn)
        ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=direct.device)
        _lib.check(lib.pvb_vote_count(direct.data_ptr(), coords.data_ptr(), hypo_pts.data_ptr(), counts.data_ptr(),
                                      tn, vn, hn, float(inlier_thresh), ws.data_ptr(), ws.numel(),
                                      _stream(direct.device)))
    return counts
