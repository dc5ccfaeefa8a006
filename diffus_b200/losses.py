"""Image-space pieces of the reference's training loops, on the device.

* :func:`ssim_loss` -- ``UltrasoundSynthesisModel.loss`` of ``notebooks/[DEMO] Train MRI to Impedance MLP - GPU.ipynb`` cell 16:
  min-max normalise the synthetic image, ``1 - piq.ssim(synth, real, data_range=1.0)``;
* :func:`masked_mse_edge_loss` -- the CPU twin's loss (``[DEMO] Train MRI to Impedance MLP.ipynb`` cell 19):
  ``mse(synth[mask], real[mask]) + 0.5 * l1(|d_x synth|[mask[:, 1:]], |d_x real|[mask[:, 1:]])``;
* :func:`log_compress` / :func:`process_rf_to_bmode` -- log compression (``[DEMO] Renderer Alternatives.ipynb`` cell 14).

All are single-image (H, W) operations on CUDA tensors, differentiable w.r.t. the synthetic image where the reference's
form is; each is one or two kernel launches with fixed-order reductions.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def ssim_loss(synth: torch.Tensor, real: torch.Tensor, *, normalize: bool = True, kernel_size: int = 11,
              kernel_sigma: float = 1.5, k1: float = 0.01, k2: float = 0.03, downsample: bool = True) -> torch.Tensor:
    """``1 - SSIM(norm(synth), real)`` with piq's ``ssim`` algorithm and defaults (data_range 1, mean reduction).

    ``normalize=True`` applies the notebook's ``(s - s.min()) / (s.max() - s.min() + 1e-8)`` first (its gradient, with
    torch's even split over tied minima / maxima, is part of the backward).  piq average-pools images whose shorter side
    is 384 pixels or more before the comparison; that branch is not implemented (B-mode images here are 256 x 256).
    """
    if downsample and max(1, round(min(synth.shape[-2:]) / 256)) > 1:
        raise NotImplementedError("piq's down-sampling of images with a side >= 384 is not implemented; pass downsample=False")
    return ops.SSIMLossFunction.apply(synth, real, bool(normalize), int(kernel_size), float(kernel_sigma), float(k1), float(k2))


def masked_mse_edge_loss(synth: torch.Tensor, real: torch.Tensor, mask: torch.Tensor, edge_weight: float = 0.5) -> torch.Tensor:
    """``mse_loss(synth[mask], real[mask]) + edge_weight * gradient_loss(synth, real, mask)`` (the CPU notebook's loss)."""
    return ops.MaskedMSEEdgeFunction.apply(synth, real, mask, float(edge_weight))


def log_compress(img: torch.Tensor) -> torch.Tensor:
    """``log1p(|img|) / max(log1p(|img|))`` -- the log-compression step of ``process_rf_to_bmode`` on an image, differentiable."""
    return ops.LogCompressFunction.apply(img)


def hilbert_kernel(n: int) -> torch.Tensor:
    """``Im(ifft(h))`` for ``scipy.signal.hilbert``'s one-sided spectrum weights ``h``: the circular-convolution kernel that turns a
    real line into the imaginary part of its analytic signal."""
    h = np.zeros(n)
    if n % 2 == 0:
        h[0] = h[n // 2] = 1
        h[1:n // 2] = 2
    else:
        h[0] = 1
        h[1:(n + 1) // 2] = 2
    return torch.tensor(np.fft.ifft(h).imag, dtype=torch.float32)


def process_rf_to_bmode(profiles: torch.Tensor) -> torch.Tensor:
    """RF-like profiles (num_rays, num_samples) -> normalised, log-compressed B-mode image (same name and arithmetic as the
    notebook's ``process_rf_to_bmode``: ``|hilbert(rf, axis=1)|`` -> ``log1p`` -> divide by the maximum); stays on the device."""
    g = hilbert_kernel(profiles.shape[1]).to(profiles.device)
    return ops.rf_to_bmode(profiles, g)
