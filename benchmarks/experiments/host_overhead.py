#!/usr/bin/env python
"""Where the host time of the reference-signature calls goes: cProfile of plot_beam_frame and of an eager
render_mse_loss + backward on a single 128 x 512 frame (config 1 / 2), plus wall and CUDA-event times per call."""
import cProfile
import io
import json
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    from diffus_b200 import UltrasoundRenderer, render_mse_loss
    from diffus_b200.phantoms import config1_pose, layered_phantom
    dev = torch.device("cuda", 0)
    vol = layered_phantom(256, seed=0).to(dev)
    src, dirs = config1_pose(256, 128)
    src, dirs = src.to(dev), dirs.to(dev)
    ren = UltrasoundRenderer(512, 1e-4)

    def fwd():
        return ren.plot_beam_frame(volume=vol, source=src, directions=dirs, plot=False, return_indices=False)

    with torch.no_grad():
        tgt = ren.plot_beam_frame(volume=vol, source=src + torch.tensor([1.5, 0.0, -1.0], device=dev), directions=dirs, plot=False,
                                  return_indices=False, sampler="trilinear")[3].unsqueeze(0)

    def step():
        s = src.clone().requires_grad_(True)
        d = dirs.clone().requires_grad_(True)
        loss = render_mse_loss(vol, s, d, tgt, 512, 1e-4)
        loss.backward()
        return s.grad

    for name, fn, n in (("plot_beam_frame", fwd, 3000), ("render_mse_loss+backward", step, 1500)):
        for _ in range(50):
            fn()
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(n):
            fn()
        host = (time.perf_counter() - t) / n          # enqueue time per call (the GPU runs behind)
        torch.cuda.synchronize()
        total = (time.perf_counter() - t) / n
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(n // 3):
            fn()
        pr.disable()
        torch.cuda.synchronize()
        sio = io.StringIO()
        pstats.Stats(pr, stream=sio).sort_stats("tottime").print_stats(22)
        print(json.dumps({"call": name, "host_us_per_call": host * 1e6, "wall_us_per_call_incl_gpu": total * 1e6}))
        print(sio.getvalue()[:6000])


if __name__ == "__main__":
    main()
