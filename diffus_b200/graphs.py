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
