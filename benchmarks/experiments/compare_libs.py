#!/usr/bin/env python
"""Kernel-development aid: run the bench's fused step with the library DIFFUS_B200_LIB points at (default: shipped) and
save loss / pose gradients, or compare two saved results.

    DIFFUS_B200_LIB=... python benchmarks/experiments/compare_libs.py run out.npz [--poses P]
    python benchmarks/experiments/compare_libs.py diff a.npz b.npz
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def run(out, poses):
    import torch
    import bench
    from diffus_b200 import PreparedVolume, ops, render_frames
    from diffus_b200._lib import SAMPLER_TRILINEAR
    dev = torch.device("cuda", 0)
    vol_h, src_h, dir_h = bench.build_scene(dev, poses, seed=1000)
    pv = PreparedVolume(vol_h.to(dev), "texture")
    s, d = src_h.to(dev), dir_h.to(dev)
    with torch.no_grad():
        tgt = render_frames(pv, s + torch.tensor([1.5, 0.0, -1.0], device=dev), d, bench.N_SAMPLES, bench.ALPHA, 0, sampler="trilinear")
    res = []
    for _ in range(3):
        loss, _, _, gs, gd = ops.render_mse_impl(pv.volume, pv.bricks, list(pv.volume.shape), s, d, tgt, bench.N_SAMPLES, 0, bench.ALPHA,
                                                 SAMPLER_TRILINEAR, False, False, True, False)
        res.append((float(loss), gs.cpu().numpy().copy(), gd.cpu().numpy().copy()))
    same = all(r[0] == res[0][0] and np.array_equal(r[1], res[0][1]) and np.array_equal(r[2], res[0][2]) for r in res)
    np.savez(out, loss=res[0][0], gs=res[0][1], gd=res[0][2], tgt_sum=float(tgt.double().sum()), repeatable=same)
    print(out, "loss", repr(res[0][0]), "repeatable", same, "target checksum", float(tgt.double().sum()))


def diff(a, b):
    A, B = np.load(a), np.load(b)
    print("loss", float(A["loss"]), float(B["loss"]), "rel", abs(float(A["loss"]) - float(B["loss"])) / abs(float(B["loss"])))
    print("target checksum equal", float(A["tgt_sum"]) == float(B["tgt_sum"]))
    for k in ("gs", "gd"):
        x, y = A[k].astype(np.float64), B[k].astype(np.float64)
        e = np.abs(x - y)
        i = np.unravel_index(e.argmax(), e.shape)
        print(k, "max abs diff", e.max(), "at", i, "values", x[i], y[i], "max |ref|", np.abs(y).max(),
              "entries differing by > 1e-4 of max:", int((e > 1e-4 * np.abs(y).max()).sum()))


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run(sys.argv[2], int(sys.argv[4]) if len(sys.argv) > 4 else 1024)
    else:
        diff(sys.argv[2], sys.argv[3])
