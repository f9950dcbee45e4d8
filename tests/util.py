"""Shared helpers for the GPU parity tests."""
import numpy as np
import torch


def field_case(tn=2000, vn=3, hn=128, seed=0, noise_deg=3.0, outliers=0.2, extent=(640, 480)):
    """Reference-layout inputs: direct [tn,vn,2], coords [tn,2] (integer pixels), idxs [hn,vn,2]."""
    rng = np.random.default_rng(seed)
    W, H = extent
    coords = np.stack([rng.integers(0, W, tn), rng.integers(0, H, tn)], axis=1).astype(np.float32)
    kp = np.stack([rng.uniform(0.1 * W, 0.9 * W, vn), rng.uniform(0.1 * H, 0.9 * H, vn)], axis=1)
    kp[-1, 0] = 1.3 * W
    ang = np.arctan2(kp[None, :, 1] - coords[:, None, 1], kp[None, :, 0] - coords[:, None, 0])
    ang = ang + rng.normal(0, np.radians(noise_deg), size=ang.shape)
    out = rng.uniform(size=ang.shape) < outliers
    ang = np.where(out, rng.uniform(0, 2 * np.pi, size=ang.shape), ang)
    direct = np.stack([np.cos(ang), np.sin(ang)], axis=-1).astype(np.float32)
    idxs = rng.integers(0, tn, size=(hn, vn, 2)).astype(np.int32)
    idxs[0, :, 1] = idxs[0, :, 0]                       # t0 == t1 -> degenerate
    return direct, coords, idxs, kp.astype(np.float32)


def cuda(*arrays):
    return [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrays]


def bits_equal(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    b = np.ascontiguousarray(b, dtype=np.float32).view(np.uint32)
    return np.array_equal(a, b)
