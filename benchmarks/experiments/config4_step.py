#!/usr/bin/env python
"""BASELINE config 4 on one GPU: a few FusedTrainer steps (for the ncu launch list), or CUDA-event timing."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sampler", default="trilinear")
    ap.add_argument("--poses", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    import bench
    from diffus_b200 import ImpedanceEstimator, PreparedVolume, render_frames
    from diffus_b200.phantoms import mri_phantom, pose_sweep
    from diffus_b200.training import FusedTrainer
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = ImpedanceEstimator(1)
    torch.manual_seed(1)
    model_tgt = ImpedanceEstimator(1)
    for m in (model, model_tgt):
        with torch.no_grad():
            m.model[4].bias.fill_(1.5)
            m.model[4].weight.mul_(0.3)
    model = model.to(dev)
    mri = (mri_phantom(256, "t2") / 1000.0).to(dev)
    s_h, d_h = pose_sweep(args.poses, 128, 256, seed=2)
    s, d = s_h.to(dev), d_h.to(dev)
    with torch.no_grad():
        z_tgt = model_tgt.to(dev).impedance_volume(mri, None, 1e6, 400.0)
        tgt = render_frames(PreparedVolume(z_tgt), s, d, 512, 1e-4, sampler=args.sampler)
    tr = FusedTrainer(model, mri, lr=1e-4, sampler=args.sampler, out_scale=1e6)
    ms = bench.timed_steps(lambda: tr.step(s, d, tgt, 512, 1e-4), 2, args.steps, torch.cuda.synchronize)
    print(json.dumps({"sampler": args.sampler, "poses": args.poses, "ms_per_step": ms, "gather": tr.gather, "loss": float(tr.loss[0])}))


if __name__ == "__main__":
    main()
