"""CUDA-graph capture of the fused pose-recovery step.

One 128 x 512 frame is ~20 us of device work but ~250 us of PyTorch eager bookkeeping (autograd engine, tensor
allocation, ctypes argument packing).  A pose-recovery loop calls the same step thousands of times with the
same shapes, so the step is captured once -- the C ABI never allocates or synchronises, which is what makes it
capturable -- and replayed with new poses written into static buffers.
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import SAMPLER_TRILINEAR
from .renderer import PreparedVolume, _resolve_start, _sampler_id


class GraphedPoseStep:
    """``loss, grad_sources, grad_directions = step(sources, directions)`` for fixed shapes, volume and targets.

    Equivalent to ``render_mse_loss(volume, sources, directions, targets, ...)`` + ``backward()``; outputs are
    static tensors that the next call overwrites.
    """

    def __init__(self, volume, targets: torch.Tensor, n_rays: int, num_samples: int, attenuation_coeff: float = 0.5,
                 start=0, sampler: str = "trilinear", shared_directions: bool = False):
        if isinstance(volume, PreparedVolume):
            bricks, vol = volume.bricks, volume.volume
        else:
            bricks, vol = None, volume.float().contiguous()
        dev = vol.device
        tgt = targets.to(torch.float32).contiguous()
        if tgt.dim() == 2:
            tgt = tgt.unsqueeze(0)
        P = tgt.shape[0]
        self.sources = torch.zeros((P, 3), dtype=torch.float32, device=dev)
        self.directions = torch.zeros((n_rays, 3) if shared_directions else (P, n_rays, 3), dtype=torch.float32, device=dev)
        self.directions[..., 1] = 1.0
        args = (vol, bricks, list(vol.shape), self.sources, self.directions, tgt, int(num_samples),
                _resolve_start(start, num_samples), float(attenuation_coeff), _sampler_id(sampler), False, False,
                _sampler_id(sampler) == SAMPLER_TRILINEAR, False)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                      # warm up outside capture (lazy module loads, attributes)
            for _ in range(2):
                ops.render_mse_impl(*args)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, _, _, self.grad_sources, self.grad_directions = ops.render_mse_impl(*args)
        self._keep = args

    def __call__(self, sources: torch.Tensor, directions: torch.Tensor):
        self.sources.copy_(sources.reshape(self.sources.shape), non_blocking=True)
        self.directions.copy_(directions.reshape(self.directions.shape), non_blocking=True)
        self.graph.replay()
        gd = self.grad_directions if self.directions.dim() == 3 else self.grad_directions.sum(0)
        return self.loss[0], self.grad_sources, gd


class GraphedFanPoseStep:
    """The same step driven by pose PARAMETERS: ``loss, g_sources, g_median, g_hint = step(sources, median, hint)``.

    The fans are generated on the device (``ops.fan_directions``) and the gradient w.r.t. the (P,R,3) directions is
    folded back onto the two (P,3) orientation vectors inside the graph, so a step moves 9 floats per pose in each
    direction instead of 3 + 3R.  Outputs are static tensors that the next call overwrites.
    """

    def __init__(self, volume, targets: torch.Tensor, n_rays: int, num_samples: int, opening_angle: float,
                 attenuation_coeff: float = 0.5, start=0):
        if isinstance(volume, PreparedVolume):
            bricks, vol = volume.bricks, volume.volume
        else:
            bricks, vol = None, volume.float().contiguous()
        dev = vol.device
        tgt = targets.to(torch.float32).contiguous()
        P = tgt.shape[0]
        # one static input block (3, P, 3) = [sources | median | hint] and one static output block
        # [g_sources | g_median | g_hint | loss]: a step is one H2D copy, one graph launch, one D2H copy
        self.poses = torch.zeros((3, P, 3), dtype=torch.float32, device=dev)
        self.sources, self.median, self.hint = self.poses[0], self.poses[1], self.poses[2]
        self.median[:, 1] = 1.0
        self.hint[:, 0] = 1.0
        self.out = torch.zeros((9 * P + 1,), dtype=torch.float32, device=dev)
        start_i = _resolve_start(start, num_samples)

        # every result is written straight into its slice of the static output block: the graph is five kernels (fans, fused
        # step, two reductions, fan backward) with no copy kernels behind them
        o_gs, o_gm, o_gh, o_loss = self.out[:3 * P], self.out[3 * P:6 * P], self.out[6 * P:9 * P], self.out[9 * P:]

        def step():
            dirs = ops.fan_directions_fwd(self.median, self.hint, opening_angle, n_rays)
            _, _, _, _, gd = ops.render_mse_impl(vol, bricks, list(vol.shape), self.sources, dirs, tgt, int(num_samples),
                                                 start_i, float(attenuation_coeff), SAMPLER_TRILINEAR, False, False, True,
                                                 False, loss_out=o_loss, grad_sources_out=o_gs)
            ops.fan_directions_bwd(self.median, self.hint, gd, opening_angle, n_rays, grad_median_out=o_gm, grad_hint_out=o_gh)

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                      # warm up outside capture
            for _ in range(2):
                step()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            step()
        self.P = P
        self.loss = self.out[9 * P]
        self.grad_sources = self.out[:3 * P].view(P, 3)
        self.grad_median = self.out[3 * P:6 * P].view(P, 3)
        self.grad_hint = self.out[6 * P:9 * P].view(P, 3)
        self._keep = (vol, bricks, tgt)

    def __call__(self, sources: torch.Tensor, median: torch.Tensor, hint: torch.Tensor):
        self.sources.copy_(sources, non_blocking=True)
        self.median.copy_(median, non_blocking=True)
        self.hint.copy_(hint, non_blocking=True)
        self.graph.replay()
        return self.loss, self.grad_sources, self.grad_median, self.grad_hint

    def packed(self, poses: torch.Tensor) -> torch.Tensor:
        """One copy in, one launch: ``poses`` is (3, P, 3) = [sources | median | hint] (e.g. pinned host memory); returns the
        static output block ``[g_sources (P,3) | g_median (P,3) | g_hint (P,3) | loss]`` as one (9P + 1) tensor."""
        self.poses.copy_(poses, non_blocking=True)
        self.graph.replay()
        return self.out
