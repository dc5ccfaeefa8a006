#!/usr/bin/env python
"""BASELINE config 5 on one GPU (512^3 volume, 512 rays x 2048 samples, poses spread over the sphere): time the fused step,
or run it once for ncu."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--poses", type=int, default=256)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--layout", default="texture")
    args = ap.parse_args()
    import bench
    rec = bench.config5_record(torch.device("cuda", 0), args.poses, args.layout) if args.iters > 1 else None
    if rec is None:
        from diffus_b200 import PreparedVolume, ops, render_frames
        from diffus_b200._lib import SAMPLER_TRILINEAR
        from diffus_b200.phantoms import layered_phantom, pose_sweep
        dev = torch.device("cuda", 0)
        pv = PreparedVolume(layered_phantom(512, 0).to(dev), args.layout)
        s_h, d_h = pose_sweep(args.poses, 512, 512, seed=3)
        s, d = s_h.to(dev), d_h.to(dev)
        with torch.no_grad():
            tgt = render_frames(pv, s + torch.tensor([1.5, 0.0, -1.0], device=dev), d, 2048, 1e-4, sampler="trilinear")
        for _ in range(2):
            ops.render_mse_impl(pv.volume, pv.bricks, [512] * 3, s, d, tgt, 2048, 0, 1e-4, SAMPLER_TRILINEAR, False, False, True, False)
        torch.cuda.synchronize()
        rec = {"ran": "2 steps"}
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
