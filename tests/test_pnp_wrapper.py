"""CPU: host-side logic of clean_pvnet_b200.uncertainty_pnp that needs no GPU -- argument checks, the Rodrigues conversion
(vs OpenCV, which the reference uses at un_pnp_utils.py:55), the options struct layout, and the reference's P3P
initialisation recipe."""
import ctypes
import os

import numpy as np
import pytest
import torch

from util import pnp_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rodrigues_matches_opencv(pvb):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    rt = rng.normal(size=(32, 6))
    rt[0, :3] = 0.0                       # identity
    rt[1, :3] = [1e-12, 0.0, -1e-12]      # tiny angle
    rt[2, :3] *= np.pi / np.linalg.norm(rt[2, :3])   # angle pi
    got = pvb.un_pnp.rodrigues(torch.from_numpy(rt)).numpy()
    for i in range(32):
        R, _ = cv2.Rodrigues(rt[i, :3])
        assert np.abs(got[i, :, :3] - R).max() < 1e-12
        assert np.array_equal(got[i, :, 3], rt[i, 3:])
    assert np.abs(np.einsum("nij,nkj->nik", got[:, :, :3], got[:, :, :3]) - np.eye(3)).max() < 1e-12


def test_batch_entry_rejects_host_tensors_and_bad_shapes(pvb):
    z = torch.zeros
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        pvb.uncertainty_pnp_batch(z(2, 9, 2), z(2, 9, 3), z(9, 3), z(3, 3), z(2, 6))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        pvb.uncertainty_pnp_batch(np.zeros((2, 9, 2)), z(2, 9, 3), z(9, 3), z(3, 3), z(2, 6))


def test_options_struct_matches_header(pvb):
    from clean_pvnet_b200 import _lib
    o = _lib.PvbPnpOptions()
    assert ctypes.sizeof(o) == 32                                   # 2 x int32 + 3 x double, no padding surprises
    assert _lib.PvbPnpOptions.function_tolerance.offset == 8 and _lib.PvbPnpOptions.parameter_tolerance.offset == 24
    hdr = open(os.path.join(ROOT, "include", "pvnet_vote_b200.h")).read()
    body = hdr[hdr.index("typedef struct pvb_pnp_options"):hdr.index("} pvb_pnp_options;")]
    names = [ln.split(";")[0].split()[-1] for ln in body.splitlines() if ";" in ln]
    assert names == [f[0] for f in _lib.PvbPnpOptions._fields_]


def test_p3p_initialisation_recipe(pvb):
    """The four best-weighted points by wxx+wxy (the reference's key, un_pnp_utils.py:25) go to OpenCV P3P; with exact
    correspondences the returned pose reprojects those four points exactly."""
    pytest.importorskip("cv2")
    uv, p3, W, K, _, true_rt = pnp_case(77, pn=9, noise=0.0)
    from clean_pvnet_b200 import uncertainty_pnp as m
    r_exp, t = m._p3p_init(p3, uv, K, W[:, 0] + W[:, 1])
    Rt = m.rodrigues(torch.from_numpy(np.concatenate([r_exp, t], 0).reshape(1, 6)))[0].numpy()
    idx = np.argsort(W[:, 0] + W[:, 1])[-4:]
    cam = p3[idx] @ Rt[:, :3].T + Rt[:, 3]
    proj = np.stack([K[0, 0] * cam[:, 0] / cam[:, 2] + K[0, 2], K[1, 1] * cam[:, 1] / cam[:, 2] + K[1, 2]], 1)
    assert np.abs(proj - uv[idx]).max() < 1e-6


def test_four_point_problem_returns_the_p3p_pose_without_a_gpu(pvb):
    """pn == 4: un_pnp_utils.py:33-37 returns the P3P pose itself; no refinement, hence no device needed."""
    pytest.importorskip("cv2")
    uv, p3, W, K, _, _ = pnp_case(78, pn=4, noise=0.0)
    Rt = pvb.un_pnp.uncertainty_pnp(uv, W, p3, K, device="cpu")
    assert Rt.shape == (3, 4) and np.isfinite(Rt).all()
    cam = p3 @ Rt[:, :3].T + Rt[:, 3]
    proj = np.stack([K[0, 0] * cam[:, 0] / cam[:, 2] + K[0, 2], K[1, 1] * cam[:, 1] / cam[:, 2] + K[1, 2]], 1)
    assert np.abs(proj - uv).max() < 1e-6
