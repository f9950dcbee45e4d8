"""Twin of `Resnet18.decode_keypoint` (lib/networks/pvnet/resnet18.py:65-76) with its front end fused:

    vertex = output['vertex'].permute(0, 2, 3, 1).view(b, h, w, vn, 2)      # consumed as a strided view, no copy
    mask   = torch.argmax(output['seg'], 1)                                 # computed inside the select kernel
    un_pnp: mean = ransac_voting_layer_v3(mask, vertex, 512, inlier_thresh=0.99)
            kpt_2d, var = estimate_voting_distribution_with_mean(mask, vertex, mean)
    else:   kpt_2d = ransac_voting_layer_v3(mask, vertex, 128, inlier_thresh=0.99, max_num=100)

The argmax never makes a round trip through HBM as a separate kernel: `pvb_decode_v3` reads the fp32 logits,
derives the foreground bitmap from the class index on the fly and writes the int64 `mask` that decode_keypoint
returns in the same pass.
"""
import torch

from . import _lib
from . import ransac_voting_gpu as _op


def decode_keypoint(output, un_pnp=True, *, seed=None, img_base=0, fused=True):
    """Adds 'mask', 'kpt_2d' (and 'var' when un_pnp) to `output` = {'seg': [b,c,h,w], 'vertex': [b,2*vn,h,w]}
    exactly like resnet18.py:65-76; returns `output`."""
    seg, ver = output["seg"], output["vertex"]
    if not (isinstance(seg, torch.Tensor) and seg.is_cuda and isinstance(ver, torch.Tensor) and ver.is_cuda):
        raise RuntimeError("seg and vertex must be CUDA tensors")
    vertex = ver.permute(0, 2, 3, 1)
    b, h, w, vn_2 = vertex.shape
    vertex = vertex.view(b, h, w, vn_2 // 2, 2)
    hn, max_num = (512, 30000) if un_pnp else (128, 100)
    if seed is None:
        seed = _op._draw_seed()
    if not fused or seg.dtype != torch.float32:
        mask = torch.argmax(seg, 1)
        mean = _op.ransac_voting_layer_v3(mask, vertex, hn, inlier_thresh=0.99, max_num=max_num, seed=seed,
                                          img_base=img_base)
    else:
        mask, mean = _decode_v3(seg, vertex, hn, 0.99, 5, max_num, seed, img_base)
    if un_pnp:
        kpt_2d, var = _op.estimate_voting_distribution_with_mean(mask, vertex, mean, seed=seed, img_base=img_base)
        output.update({"mask": mask, "kpt_2d": kpt_2d, "var": var})
    else:
        output.update({"mask": mask, "kpt_2d": mean})
    return output


def _decode_v3(seg, vertex, hn, inlier_thresh, min_num, max_num, seed, img_base):
    lib = _lib.load()
    dev = vertex.device
    if vertex.dtype != torch.float32:
        vertex = vertex.float()
    B, H, W, K, _ = vertex.shape
    C = seg.shape[1]
    with torch.cuda.device(dev):
        d = _op._make_desc(torch.empty((0, 1, 1), dtype=torch.int64), vertex, hn, inlier_thresh, min_num, max_num,
                           _lib.PVB_SELECT_BYTE, seed, img_base, None)
        d.mask_stride[0], d.mask_stride[1], d.mask_stride[2] = seg.stride(0), seg.stride(2), seg.stride(3)
        nbytes = lib.pvb_workspace_bytes(d)
        ws = _op._workspace(dev, nbytes)
        mask = torch.empty((B, H, W), dtype=torch.int64, device=dev)
        out = torch.empty((B, K, 2), dtype=torch.float32, device=dev)
        if B:
            _lib.check(lib.pvb_decode_v3(d, seg.data_ptr(), C, seg.stride(1), mask.data_ptr(), vertex.data_ptr(), None,
                                         None, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                         torch.cuda.current_stream(dev).cuda_stream))
    return mask, out


def uncertainty_pnp_weights(var):
    """inv(sqrtm(var)) per keypoint packed as (wxx, wxy, wyy) -- the `weights_2d` argument of
    un_pnp_utils.uncertainty_pnp -- computed on the GPU in closed form instead of the per-keypoint
    scipy.linalg.sqrtm / np.linalg.inv loop of lib/evaluators/linemod/pvnet.py:118-130.
    var: CUDA float tensor [..., 2, 2] -> float32 [..., 3]."""
    if not isinstance(var, torch.Tensor) or not var.is_cuda or var.shape[-2:] != (2, 2):
        raise RuntimeError("var must be a CUDA tensor [...,2,2]")
    lib = _lib.load()
    v = var.float().contiguous()
    out = torch.empty(tuple(v.shape[:-2]) + (3,), dtype=torch.float32, device=v.device)
    n = v.numel() // 4
    with torch.cuda.device(v.device):
        if n:
            _lib.check(lib.pvb_uncertainty_weights(v.data_ptr(), out.data_ptr(), n,
                                                   torch.cuda.current_stream(v.device).cuda_stream))
    return out
