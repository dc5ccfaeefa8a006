"""diffus_b200 -- B200-native DiffUS B-mode renderer hot path.

Drop-in for ``UltrasoundRenderer.plot_beam_frame`` (+ ``generate_cone_directions`` and
``ImpedanceEstimator``) of gduguey/DiffUS over hand-written sm_100a kernels behind a C ABI
(``include/diffus_b200.h``).  Importing the package is cheap and works without a GPU; the
first op call loads ``libdiffus_b200.so`` and fails loudly if it has not been built.
"""
from .cone import generate_cone_directions
from .impedance import ImpedanceEstimator
from .ops import fan_directions
from .renderer import (PreparedVolume, UltrasoundRenderer, compute_echo_traces, compute_gaussian_pulse, gaussian_pulse, custom_nearest_sampler, differentiable_splat,
                       propagate_full_rays_batched, render_frames, render_mse_loss, rotate_around_apex)

__all__ = ["UltrasoundRenderer", "render_frames", "render_mse_loss", "PreparedVolume", "compute_echo_traces", "compute_gaussian_pulse", "gaussian_pulse",
           "propagate_full_rays_batched", "custom_nearest_sampler", "differentiable_splat", "rotate_around_apex", "generate_cone_directions",
           "fan_directions", "ImpedanceEstimator"]
