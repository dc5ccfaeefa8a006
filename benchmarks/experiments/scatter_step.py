"""The render + d loss / d volume scatter launch of the MLP-training step (BASELINE config 4) on its own:

    python benchmarks/experiments/scatter_step.py [--sampler trilinear|nearest] [--poses 4096] [--iters 5] [--no-grad]

Times `render_bwd_kernel<..., vol_grad, LOSS_MSE>` (fused forward + MSE + backward + scatter into a BRICK-layout gradient
volume) with CUDA events; `--no-grad` runs the same launch without the volume gradient (the render part alone).  Used for
the A/B numbers in profiles/ and as the short command behind the ncu captures of the scatter kernel.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from diffus_b200 import PreparedVolume, ops, render_frames  # noqa: E402
from diffus_b200._lib import SAMPLER_NEAREST, SAMPLER_TRILINEAR  # noqa: E402
from diffus_b200.phantoms import intensity_to_impedance, mri_phantom, pose_sweep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sampler", default="trilinear", choices=["trilinear", "nearest"])
    ap.add_argument("--poses", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--no-grad", action="store_true")
    ap.add_argument("--check", action="store_true", help="compare the gradient volume with a float64 accumulation of per-pose launches")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    vol = intensity_to_impedance(mri_phantom(256, "t2")).to(dev)
    pv = PreparedVolume(vol, "brick")
    s, d = pose_sweep(args.poses, 128, 256, seed=2)
    s, d = s.to(dev), d.to(dev)
    sid = SAMPLER_TRILINEAR if args.sampler == "trilinear" else SAMPLER_NEAREST
    with torch.no_grad():
        tgt = render_frames(pv, s + 0.5, d, 512, 1e-4, sampler=args.sampler)

    def step():
        return ops.render_mse_impl(vol, pv.bricks, [256] * 3, s, d, tgt, 512, 0, 1e-4, sid, False, not args.no_grad, False,
                                   False, keep_brick_grad=True)
    for _ in range(2):
        out = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    line = {"sampler": args.sampler, "poses": args.poses, "volume_grad": not args.no_grad, "ms": ms,
            "ms_per_4096_poses": ms * 4096 / args.poses, "lib": os.environ.get("DIFFUS_B200_LIB", "shipped"),
            "loss": float(out[0]), "grad_abs_sum": float(out[2].double().abs().sum()) if not args.no_grad else None}
    if args.check and not args.no_grad:
        ref = torch.zeros_like(out[2], dtype=torch.float64)
        n = tgt.numel()
        for p0 in range(0, args.poses, 64):          # small launches, accumulated in float64: a rounding-independent reference
            sl = slice(p0, min(p0 + 64, args.poses))
            g = ops.render_mse_impl(vol, pv.bricks, [256] * 3, s[sl], d[sl], tgt[sl], 512, 0, 1e-4, sid, False, True, False, False,
                                    keep_brick_grad=True)[2]
            ref += g.double() * (tgt[sl].numel() / n)
        err = (out[2].double() - ref).abs().max().item() / ref.abs().max().item()
        line["max_err_vs_chunked_f64_over_max"] = err
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
