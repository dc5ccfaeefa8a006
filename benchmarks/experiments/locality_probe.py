#!/usr/bin/env python
"""Kernel-development probe: how much of the fused pose step's time is gather locality?  Times the step on the bench scene
(1024 different poses) and on degenerate scenes whose poses repeat (every CTA then walks texels its neighbours just fetched)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    import bench
    from diffus_b200 import PreparedVolume, ops, render_frames
    from diffus_b200._lib import SAMPLER_TRILINEAR
    dev = torch.device("cuda", 0)
    vol_h, src_h, dir_h = bench.build_scene(dev, 1024, seed=1000)
    pv = PreparedVolume(vol_h.to(dev), "texture")
    for name, rep in (("1024 distinct poses", 1), ("each pose 8 times in a row", 8), ("each pose 64 times in a row", 64), ("one pose 1024 times", 1024)):
        idx = (torch.arange(1024) // rep) * rep
        s, d = src_h[idx].contiguous().to(dev), dir_h[idx].contiguous().to(dev)
        with torch.no_grad():
            tgt = render_frames(pv, s + torch.tensor([1.5, 0.0, -1.0], device=dev), d, 512, 1e-4, 0, sampler="trilinear")

        def step():
            ops.render_mse_impl(pv.volume, pv.bricks, list(pv.volume.shape), s, d, tgt, 512, 0, 1e-4, SAMPLER_TRILINEAR, False, False, True, False)
        ms = bench.timed_steps(step, 5, 50, torch.cuda.synchronize)
        print(json.dumps({"scene": name, "ms": ms}), flush=True)


if __name__ == "__main__":
    main()
