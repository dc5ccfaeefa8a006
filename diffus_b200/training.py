"""MRI -> impedance MLP training fused with rendering (BASELINE config 4; reference notebooks
``[DEMO] Train MRI to Impedance MLP*.ipynb``, ``ImpedanceLearner.training_forward`` + ``train_step``).

One step, everything in the brick layout the gathers use (no layout conversions inside the step):

    Z_bricks  = out_scale * MLP(mri_bricks)           tcgen05 kernel, written once
    loss, dZ  = fused render + MSE + backward          dZ scattered with red.add into brick-local lines
    dWeights  = MLP backward from (mri_bricks, dZ)     atomic-free block partials, fixed-order reduction
    all-reduce of the 1 153 weight gradients (+ shared-pose gradients) across ranks, one flat buffer
"""
from __future__ import annotations

from typing import Optional

import torch

from . import distributed as dist_utils
from . import ops
from ._lib import SAMPLER_TRILINEAR
from .impedance import ImpedanceEstimator, pack_params
from .renderer import _canon_pose, _resolve_start, _sampler_id


class TrainingVolume:
    """An MRI volume (and optional mask) copied once into the brick layout of the gathers."""

    def __init__(self, mri: torch.Tensor, mask: Optional[torch.Tensor] = None):
        if mri.dim() != 3:
            raise ValueError("mri must be (D,H,W)")
        self.dims = list(mri.shape)
        self.mri_bricks = ops.to_bricks(mri.float())
        self.mask_bricks = None if mask is None else (ops.to_bricks(mask.float()) > 0.5).to(torch.uint8)


class _MLPRenderMSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, params, tv, sources, directions, targets, n_samples, start, alpha, sampler, product_f32,
                out_scale, fill):
        z_bricks = ops.mlp_fwd_impl(params.detach(), tv.mri_bricks, tv.mask_bricks, out_scale, fill)
        need_pose = (ctx.needs_input_grad[2] or ctx.needs_input_grad[3]) and sampler == SAMPLER_TRILINEAR
        loss, _, gz, gsrc, gdir = ops.render_mse_impl(z_bricks, z_bricks, tv.dims, sources.detach(), directions.detach(),
                                                      targets, n_samples, start, alpha, sampler, product_f32,
                                                      ctx.needs_input_grad[0], need_pose, False, keep_brick_grad=True)
        ctx.save_for_backward(params, gz, gsrc, gdir)
        ctx.tv, ctx.out_scale = tv, out_scale
        ctx.flags = (need_pose, sources.dtype, directions.dtype, directions.dim())
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gloss):
        params, gz, gsrc, gdir = ctx.saved_tensors
        need_pose, sdt, ddt, ddim = ctx.flags
        gp = gs = gd = None
        if ctx.needs_input_grad[0]:
            gp = ops.mlp_bwd_impl(params.detach(), ctx.tv.mri_bricks, ctx.tv.mask_bricks, gz, ctx.out_scale) * gloss
        if need_pose and ctx.needs_input_grad[2]:
            gs = (gsrc * gloss).to(sdt)
        if need_pose and ctx.needs_input_grad[3]:
            gd = ((gdir if ddim == 3 else gdir.sum(0)) * gloss).to(ddt)
        return gp, None, gs, gd, None, None, None, None, None, None, None, None


def mlp_render_mse_loss(model: ImpedanceEstimator, mri, sources: torch.Tensor, directions: torch.Tensor,
                        targets: torch.Tensor, num_samples: int, attenuation_coeff: float = 0.5, start=0, *,
                        sampler: str = "trilinear", out_scale: float = 1.0, mask: Optional[torch.Tensor] = None,
                        fill: float = 400.0) -> torch.Tensor:
    """MSE between frames rendered from ``out_scale * MLP(mri)`` and ``targets``.

    Differentiable in the MLP weights and (trilinear) the poses.  ``mri`` is a (D,H,W) CUDA tensor or a
    :class:`TrainingVolume` prepared once (saves the per-step brick copy of the input).
    """
    tv = mri if isinstance(mri, TrainingVolume) else TrainingVolume(mri, mask)
    src, dirs, product_f32 = _canon_pose(sources, directions, tv.mri_bricks.device)
    if src.dim() == 1:
        src = src.unsqueeze(0)
    tgt = targets.to(torch.float32).contiguous()
    if tgt.dim() == 2:
        tgt = tgt.unsqueeze(0)
    return _MLPRenderMSE.apply(pack_params(model), tv, src.contiguous(), dirs.contiguous(), tgt, int(num_samples),
                               _resolve_start(start, num_samples), float(attenuation_coeff), _sampler_id(sampler),
                               product_f32, float(out_scale), float(fill))


def train_step(model: ImpedanceEstimator, optimizer: torch.optim.Optimizer, mri, sources: torch.Tensor,
               directions: torch.Tensor, targets: torch.Tensor, num_samples: int, attenuation_coeff: float = 0.5,
               start=0, **kw) -> torch.Tensor:
    """zero_grad -> fused loss -> backward -> all-reduce(mean) of the weight gradients -> optimizer.step()."""
    optimizer.zero_grad(set_to_none=True)
    loss = mlp_render_mse_loss(model, mri, sources, directions, targets, num_samples, attenuation_coeff, start, **kw)
    loss.backward()
    dist_utils.allreduce_module_grads(model, average=True)
    optimizer.step()
    return loss.detach()
