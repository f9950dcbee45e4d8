#!/usr/bin/env python
"""Generates the golden vectors in tests/golden/*.npz by running the UNMODIFIED reference
(oracle/_ref: clean-pvnet's CUDA extension + its Python operator, built by oracle/build_ref.py)
on seeded inputs.  Needs a GPU:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'  &&  cp gpurun_out/golden/*.npz tests/golden/

The reference has no golden vectors of its own; these are the pin for the CPU
oracle (tests/test_golden.py, CPU) and for the CUDA path (tests/test_gpu_golden.py).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from refload import load_reference  # noqa: E402
from util import field_case  # noqa: E402


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    ext, gpu = load_reference()
    dev = "cuda"
    # ---- kernel level ---------------------------------------------------------------------
    direct, coords, idxs, kp = field_case(tn=600, vn=3, hn=48, seed=21)
    d, c, i = [torch.from_numpy(a).to(dev) for a in (direct, coords, idxs)]
    hyp = ext.generate_hypothesis(d, c, i)
    counts = {}
    for th in (0.99, 0.999):
        inl = torch.zeros((48, 3, 600), dtype=torch.uint8, device=dev)
        ext.voting_for_hypothesis(d, c, hyp, inl, th)
        counts[th] = inl.sum(dim=2, dtype=torch.int32).cpu().numpy()
    inl99 = torch.zeros((48, 3, 600), dtype=torch.uint8, device=dev)
    ext.voting_for_hypothesis(d, c, hyp, inl99, 0.99)
    hyp_vp = ext.generate_hypothesis_vanishing_point(d, c, i)
    inl_vp = torch.zeros((48, 3, 600), dtype=torch.uint8, device=dev)
    ext.voting_for_hypothesis_vanishing_point(d, c, hyp_vp, inl_vp, 0.999)
    np.savez_compressed(os.path.join(out_dir, "kernels.npz"), direct=direct, coords=coords, idxs=idxs,
                        hyp=hyp.cpu().numpy(), counts_099=counts[0.99], counts_0999=counts[0.999],
                        inliers_099_packed=np.packbits(inl99.cpu().numpy(), axis=2),
                        hyp_vp=hyp_vp.cpu().numpy(), inliers_vp_packed=np.packbits(inl_vp.cpu().numpy(), axis=2))
    # ---- operator level -------------------------------------------------------------------
    from clean_pvnet_b200 import ransac_voting_gpu as mine
    from clean_pvnet_b200 import synth
    mask, vertex, kpt = synth.make_inputs("tiny", device=dev, seed=4321)
    B, H, W, K, _ = vertex.shape
    for name, hn, max_num in (("v3_plain", 32, 30000), ("v3_thinned", 32, 300)):
        torch.manual_seed(11)
        want = gpu.ransac_voting_layer_v3(mask, vertex, hn, inlier_thresh=0.99, max_num=max_num)
        torch.manual_seed(11)
        idx_t, sel_t = mine._torch_rng_draws(mask, hn, K, 5, max_num, 0)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), mask=mask.cpu().numpy().astype(np.uint8),
                            vertex=vertex.cpu().numpy(), hn=hn, max_num=max_num, thresh=np.float32(0.99),
                            idxs=idx_t.cpu().numpy(),
                            selection=(sel_t.cpu().numpy() if sel_t is not None else np.zeros(0, np.float32)),
                            kpt=want.cpu().numpy())
    mean = gpu.ransac_voting_layer_v3(mask, vertex, 32, inlier_thresh=0.99)
    torch.manual_seed(12)
    _, cov = gpu.estimate_voting_distribution_with_mean(mask, vertex, mean.clone(), round_hyp_num=32, min_hyp_num=128)
    torch.manual_seed(12)
    idx_t, sel_t = mine._torch_rng_draws(mask, 32, K, 5, 30000, 1, rounds=4)
    np.savez_compressed(os.path.join(out_dir, "dist.npz"), mask=mask.cpu().numpy().astype(np.uint8),
                        vertex=vertex.cpu().numpy(), mean=mean.cpu().numpy(), idxs=idx_t.cpu().numpy(),
                        cov=cov.cpu().numpy(), round_hyp_num=32, min_hyp_num=128)
    print("golden vectors written to", out_dir, os.listdir(out_dir))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE))
