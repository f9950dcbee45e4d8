#!/usr/bin/env python
"""One (or a few) ransac_voting_layer_v3 steps on BASELINE cfg2 for profiling under ncu.

    ncu --set full --clock-control none --import-source on -k regex:vote_kernel -c 1 -o gpurun_out/vote python tools/prof_step.py
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pvb --csv --log-file gpurun_out/launches.csv python tools/prof_step.py --steps 3
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import clean_pvnet_b200 as pvb  # noqa: E402
from clean_pvnet_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--cfg", default="cfg2")
ap.add_argument("--hn", type=int, default=512)
ap.add_argument("--B", type=int, default=None)
ap.add_argument("--dist", action="store_true", help="also run estimate_voting_distribution_with_mean")
ap.add_argument("--gather", type=int, default=0, help="gather_mode of pvb_set_tuning")
ap.add_argument("--variant", type=int, default=0)
args = ap.parse_args()
from clean_pvnet_b200 import _lib  # noqa: E402
_lib.check(_lib.load().pvb_set_tuning(args.gather, args.variant))
mask, vertex, _ = synth.make_inputs(args.cfg, device="cuda", seed=1236, B=args.B)
for i in range(args.steps):
    out = pvb.ransac_voting_layer_v3(mask, vertex, args.hn, inlier_thresh=0.99, seed=1000 + i)
    if args.dist:
        pvb.estimate_voting_distribution_with_mean(mask, vertex, out, seed=2000 + i)
torch.cuda.synchronize()
print("ok", out[0, 0].tolist())
