"""MRI -> acoustic impedance MLP -- drop-in for ``ImpedanceEstimator`` (reference ``src/impedance.py:6-54``).

``Linear(1,32)-ReLU-Linear(32,32)-ReLU-Linear(32,1)`` (1 153 parameters).  The module keeps
the reference's ``nn.Sequential`` parameter names (``model.0.weight`` ...) so state dicts are
interchangeable; on CUDA inputs with ``input_dim == 1`` the forward and the weight
gradient run in the fused volume kernels of ``csrc/mlp_kernels.cu``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


def pack_params(module: "ImpedanceEstimator") -> torch.Tensor:
    """[W1 b1 W2 b2 W3 b3] flattened in ``nn.Linear`` (out, in) layout -- differentiable ``cat``."""
    seq = module.model
    return torch.cat([seq[0].weight.reshape(-1), seq[0].bias, seq[2].weight.reshape(-1), seq[2].bias,
                      seq[4].weight.reshape(-1), seq[4].bias])


class ImpedanceEstimator(nn.Module):
    """MLP for estimating acoustic impedance from normalized intensity values."""

    def __init__(self, input_dim: int = 1):
        super().__init__()
        self.input_dim = input_dim
        self.model = nn.Sequential(
            nn.Linear(input_dim, 32), nn.ReLU(),
            nn.Linear(32, 32), nn.ReLU(),
            nn.Linear(32, 1)
        )

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.input_dim == 1 and x.is_cuda and x.shape[-1] == 1 and x.dtype == torch.float32:
            out = ops.mlp_fwd(pack_params(self), x.reshape(-1), None, 1.0, 0.0)
            return out.reshape(x.shape)
        if x.is_cuda:
            raise NotImplementedError("the fused MLP kernels cover input_dim == 1 float32 (every use in the reference)")
        from ._lib import DiffusError
        raise DiffusError("ImpedanceEstimator runs only on CUDA tensors (sm_100a kernels); there is no CPU fallback -- "
                          "move the model and its inputs to the GPU")

    def impedance_volume(self, volume: torch.Tensor, mask: torch.Tensor = None, out_scale: float = 1.0,
                         fill: float = 0.0) -> torch.Tensor:
        """Evaluate the MLP on every voxel of a CUDA volume in one fused pass (differentiable in the weights).

        ``mask`` (bool/uint8, same shape) selects voxels; the others get ``fill``.
        """
        out = ops.mlp_fwd(pack_params(self), volume.reshape(-1), None if mask is None else mask.reshape(-1),
                          float(out_scale), float(fill))
        return out.reshape(volume.shape)

    @classmethod
    def train_model(cls, X: torch.Tensor, y: torch.Tensor, input_dim: int = 1, lr: float = 1e-3,
                    epochs: int = 5000) -> "ImpedanceEstimator":
        """Supervised fit on paired (intensity, impedance) rows (reference ``:19-37``): Adam + MSE, full batch."""
        model = cls(input_dim).to(X.device)
        optimizer = torch.optim.Adam(model.parameters(), lr=lr)
        loss_fn = nn.MSELoss()
        for _ in range(epochs):
            optimizer.zero_grad()
            loss = loss_fn(model(X), y)
            loss.backward()
            optimizer.step()
        return model

    @staticmethod
    def compute_impedance_volume(volume: torch.Tensor, model: "ImpedanceEstimator", threshold: float = 50,
                                 mask: torch.Tensor = None) -> torch.Tensor:
        """Full impedance volume from a trained model (reference ``:39-54``), all on the device.

        ``create_brain_mask`` (threshold + 2 dilations + 2 erosions) -> ``zscore_normalize`` inside the mask ->
        MLP x 1e6 on masked voxels -> 400.0 (air) elsewhere.  Pass ``mask`` to supply your own.
        """
        from .utils import create_brain_mask, zscore_normalize
        if mask is None:
            mask = create_brain_mask(volume, threshold)
        vol_norm = zscore_normalize(volume, mask)
        with torch.no_grad():
            return model.impedance_volume(vol_norm, mask, out_scale=1e6, fill=400.0).to(volume.dtype)
