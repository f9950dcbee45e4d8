"""Synthetic masks / vertex fields of the shapes BASELINE.json names.

Vertex vectors follow the reference's ground-truth builder `compute_vertex`
(lib/utils/pvnet/pvnet_data_utils.py:30-44): unit vector from the pixel (x=col, y=row) towards
the keypoint; on top of that angular noise, a fraction of uniformly random directions
(outliers) on the foreground, and random unit vectors on the background (network garbage that
the mask must filter out).  Works on CPU or CUDA tensors; deterministic per (seed, device type).
"""
import math

import torch

CONFIGS = {
    # name: B, H, W, K, hn, fill (lo, hi), mask kind
    "cfg1": dict(B=1, H=128, W=128, K=1, hn=64, fill=(0.30, 0.30), kind="blob", random_field=True),
    "cfg2": dict(B=16, H=480, W=640, K=9, hn=512, fill=(0.30, 0.30), kind="blob"),
    "cfg3": dict(B=64, H=480, W=640, K=9, hn=1024, fill=(0.05, 0.15), kind="fragmented"),
    "cfg4": dict(B=128, H=720, W=540, K=17, hn=512, fill=(0.30, 0.30), kind="blob"),
    "cfg5": dict(B=256, H=640, W=640, K=9, hn=512, fill=(0.01, 0.80), kind="blob"),
    "tiny": dict(B=2, H=48, W=64, K=3, hn=32, fill=(0.30, 0.30), kind="blob"),
    "small": dict(B=3, H=96, W=128, K=4, hn=64, fill=(0.25, 0.35), kind="blob"),
}


def _blob_mask(H, W, fill, g, device):
    """Filled random ellipse whose area is fill*H*W (clipped to the image)."""
    yy = torch.arange(H, device=device, dtype=torch.float32)[:, None]
    xx = torch.arange(W, device=device, dtype=torch.float32)[None, :]
    area = fill * H * W
    aspect = 0.6 + 0.8 * torch.rand((), generator=g, device=device).item()
    a = math.sqrt(area / math.pi * aspect)
    b = area / (math.pi * a)
    a, b = min(a, 0.49 * W), min(b, 0.49 * H)
    if math.pi * a * b < area:   # very large fill: grow into a super-ellipse box
        a, b = 0.5 * W * math.sqrt(fill) * 1.13, 0.5 * H * math.sqrt(fill) * 1.13
    cx = a + (W - 2 * a) * torch.rand((), generator=g, device=device).item() if W > 2 * a else W / 2
    cy = b + (H - 2 * b) * torch.rand((), generator=g, device=device).item() if H > 2 * b else H / 2
    th = math.pi * torch.rand((), generator=g, device=device).item() * 0.25
    dx, dy = xx - cx, yy - cy
    u = dx * math.cos(th) + dy * math.sin(th)
    v = -dx * math.sin(th) + dy * math.cos(th)
    return ((u / a) ** 2 + (v / b) ** 2) <= 1.0


def _fragmented_mask(H, W, fill, g, device):
    """Union of 8-32 discs with Bernoulli(0.7) pixel dropout, total fill ~ `fill`."""
    yy = torch.arange(H, device=device, dtype=torch.float32)[:, None]
    xx = torch.arange(W, device=device, dtype=torch.float32)[None, :]
    n = int(torch.randint(8, 33, (), generator=g, device=device).item())
    r = math.sqrt(fill * H * W / 0.7 / (math.pi * n)) * 1.1
    m = torch.zeros((H, W), dtype=torch.bool, device=device)
    c = torch.rand((n, 2), generator=g, device=device)
    for i in range(n):
        cx, cy = float(c[i, 0]) * W, float(c[i, 1]) * H
        m |= ((xx - cx) ** 2 + (yy - cy) ** 2) <= r * r
    keep = torch.rand((H, W), generator=g, device=device) < 0.7
    return m & keep


def make_inputs(cfg, device="cpu", seed=1234, noise_deg=3.0, outlier_frac=0.20, mask_dtype=torch.int64,
                layout="interleaved", B=None):
    """Returns (mask [B,H,W] mask_dtype {0,1}, vertex [B,H,W,K,2] float32, keypoints [B,K,2]).

    layout "interleaved": vertex is contiguous [B,H,W,K,2] (what BASELINE.json names);
    layout "planar": vertex is the permuted view of a contiguous [B,2K,H,W] tensor, i.e. exactly
    what decode_keypoint passes (resnet18.py:66-68)."""
    c = dict(CONFIGS[cfg]) if isinstance(cfg, str) else dict(cfg)
    if B is not None:
        c["B"] = B
    B_, H, W, K = c["B"], c["H"], c["W"], c["K"]
    device = torch.device(device)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    masks, verts, kpts = [], [], []
    yy = torch.arange(H, device=device, dtype=torch.float32)[:, None, None]
    xx = torch.arange(W, device=device, dtype=torch.float32)[None, :, None]
    for _ in range(B_):
        lo, hi = c["fill"]
        fill = lo + (hi - lo) * torch.rand((), generator=g, device=device).item()
        m = _blob_mask(H, W, fill, g, device) if c["kind"] == "blob" else _fragmented_mask(H, W, fill, g, device)
        kp = torch.rand((K, 2), generator=g, device=device)
        kp = torch.stack([(0.1 + 0.8 * kp[:, 0]) * W, (0.1 + 0.8 * kp[:, 1]) * H], dim=1)
        if K > 1:   # one keypoint outside the image: voting must extrapolate
            kp[K - 1, 0] = W * 1.3
            kp[K - 1, 1] = H * (0.2 + 0.6 * torch.rand((), generator=g, device=device).item())
        ang = torch.atan2(kp[None, None, :, 1] - yy, kp[None, None, :, 0] - xx)          # [H,W,K]
        ang = ang + torch.randn((H, W, K), generator=g, device=device) * math.radians(noise_deg)
        rnd = torch.rand((H, W, K), generator=g, device=device) * (2 * math.pi)
        is_out = torch.rand((H, W, K), generator=g, device=device) < outlier_frac
        if c.get("random_field"):
            is_out = torch.ones_like(is_out)
        ang = torch.where(is_out | ~m[:, :, None], rnd, ang)
        v = torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1)                          # [H,W,K,2]
        masks.append(m)
        verts.append(v)
        kpts.append(kp)
    mask = torch.stack(masks).to(mask_dtype)
    vertex = torch.stack(verts).contiguous()
    if layout == "planar":
        nchw = vertex.view(B_, H, W, 2 * K).permute(0, 3, 1, 2).contiguous()               # [B,2K,H,W]
        vertex = nchw.permute(0, 2, 3, 1).view(B_, H, W, K, 2)
    elif layout != "interleaved":
        raise ValueError(layout)
    return mask, vertex, torch.stack(kpts)
