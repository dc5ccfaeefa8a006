"""MRI -> impedance MLP training fused with rendering (BASELINE config 4; reference notebooks
``[DEMO] Train MRI to Impedance MLP*.ipynb``, ``ImpedanceLearner.training_forward`` + ``train_step``).

One step:  Z = out_scale * MLP(mri) over the whole volume (fused kernel, written once)
        -> brick copy of Z for the gathers
        -> fused render + MSE + backward over the local poses (d loss/dZ scattered with red.add)
        -> MLP weight gradient from the dZ volume (atomic-free block partials, fixed-order reduction)
        -> one flat all-reduce of the 1 153 weight gradients (+ shared-pose gradients) across ranks.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import distributed as dist_utils
from .impedance import ImpedanceEstimator
from .renderer import PreparedVolume, render_mse_loss


def mlp_render_mse_loss(model: ImpedanceEstimator, mri: torch.Tensor, sources: torch.Tensor,
                        directions: torch.Tensor, targets: torch.Tensor, num_samples: int,
                        attenuation_coeff: float = 0.5, start=0, *, sampler: str = "trilinear",
                        out_scale: float = 1.0, mask: Optional[torch.Tensor] = None, fill: float = 400.0,
                        bricks: bool = True) -> torch.Tensor:
    """MSE between frames rendered from ``MLP(mri)`` and ``targets``; differentiable in the MLP weights and poses."""
    Z = model.impedance_volume(mri, mask, out_scale=out_scale, fill=fill)
    vol = PreparedVolume(Z) if bricks else Z
    return render_mse_loss(vol, sources, directions, targets, num_samples, attenuation_coeff, start, sampler=sampler)


def train_step(model: ImpedanceEstimator, optimizer: torch.optim.Optimizer, mri: torch.Tensor,
               sources: torch.Tensor, directions: torch.Tensor, targets: torch.Tensor, num_samples: int,
               attenuation_coeff: float = 0.5, start=0, **kw) -> torch.Tensor:
    """zero_grad -> fused loss -> backward -> all-reduce(mean) of the weight gradients -> optimizer.step()."""
    optimizer.zero_grad(set_to_none=True)
    loss = mlp_render_mse_loss(model, mri, sources, directions, targets, num_samples, attenuation_coeff, start, **kw)
    loss.backward()
    dist_utils.allreduce_module_grads(model, average=True)
    optimizer.step()
    return loss.detach()
