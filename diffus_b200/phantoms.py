"""Synthetic inputs for the benchmark configurations.

No dataset volumes ship with the reference (``.gitignore:1-7`` there), so every
measurement and parity test runs on seeded synthetic volumes of the shapes the reference's
notebooks use.  Everything is generated on the CPU with ``torch.Generator`` so that the
same seed gives the same tensor on the build container and on the GPU box.
"""
from __future__ import annotations

import math

import torch

# tissue impedances [Rayl] quoted by the reference (USPhysics.md:52-59 and the tissue
# table of notebooks/[TEST] Different orientations.ipynb cell 5)
LAYER_IMPEDANCE = (1.52e6, 1.34e6, 1.68e6, 1.60e6, 1.50e6, 1.67e6, 1.60e6, 1.38e6)
# T1 intensities of notebooks/[DEMO] REUBEN DATA 46.ipynb cell 1
T1_INTENSITY = {"air": 0.0, "fat": 260.0, "wm": 780.0, "gm": 920.0, "csf": 2500.0}


def layered_phantom(n: int = 256, seed: int = 0, noise: float = 0.01) -> torch.Tensor:
    """Config 1/2/5 volume: wavy tissue layers stacked along axis 1, 1 % multiplicative noise."""
    g = torch.Generator().manual_seed(seed)
    s = n / 256.0
    i = torch.arange(n, dtype=torch.float32)
    x, y, z = i.view(n, 1, 1), i.view(1, n, 1), i.view(1, 1, n)
    depth = y + 10 * s * torch.sin(4 * math.pi * x / n) + 6 * s * torch.cos(6 * math.pi * z / n)
    layer = torch.clamp(torch.floor(depth / (32 * s)), 0, 7).long()
    vol = torch.tensor(LAYER_IMPEDANCE, dtype=torch.float32)[layer]
    vol = vol * (1 + noise * torch.randn(vol.shape, generator=g))
    return vol.contiguous()


def mri_phantom(n: int = 256, kind: str = "t1", seed: int = 0) -> torch.Tensor:
    """Config 3/4 volume: ellipsoid head with CSF shell and a tumour box, MRI intensities."""
    g = torch.Generator().manual_seed(seed)
    t = torch.linspace(-1, 1, n)
    x, y, z = t.view(n, 1, 1), t.view(1, n, 1), t.view(1, 1, n)
    brain = (x / 0.8) ** 2 + (y / 0.95) ** 2 + (z / 0.85) ** 2 <= 1.0
    shell = ((x / 0.88) ** 2 + (y / 1.05) ** 2 + (z / 0.93) ** 2 <= 1.0) & ~brain
    white = (x / 0.55) ** 2 + (y / 0.7) ** 2 + (z / 0.6) ** 2 <= 1.0
    tumour = (x.abs() < 0.2) & (y.abs() < 0.3) & (z.abs() < 0.25)
    v = T1_INTENSITY
    if kind == "t1":
        gm, wm, csf, tum = v["gm"], v["wm"], v["csf"], v["fat"]
    elif kind == "t2":                      # CSF brightest, white/grey swapped
        gm, wm, csf, tum = v["wm"], v["fat"], v["csf"] * 1.2, v["gm"]
    else:
        raise ValueError(kind)
    vol = torch.zeros((n, n, n), dtype=torch.float32)
    vol = torch.where(shell, torch.tensor(csf), vol)
    vol = torch.where(brain, torch.tensor(gm), vol)
    vol = torch.where(white, torch.tensor(wm), vol)
    vol = torch.where(tumour & brain, torch.tensor(tum), vol)
    vol = vol * (1 + 0.02 * torch.randn(vol.shape, generator=g))
    return vol.contiguous()


def intensity_to_impedance(mri: torch.Tensor) -> torch.Tensor:
    """A smooth strictly positive stand-in for a trained MLP.

    Background (intensity 0) maps to coupling gel / water (1.48e6 Rayl) rather than air: a
    probe is in contact with tissue, and an air gap (Z = 400, |r| ~ 1) makes the layered-medium
    echo P01/P11 resonate to values of 1e3 and more, which says nothing about a renderer.
    Tissue lands in 1.5e6 .. 1.7e6, so |r| stays below ~0.07 as in the reference's tissue tables.
    """
    u = mri / 2500.0
    return (1.48e6 + 2.2e5 * torch.tanh(2 * u)).contiguous()


def fan_directions(median: torch.Tensor, normal_hint: torch.Tensor, opening_angle: float,
                   n_rays: int) -> torch.Tensor:
    """Fans in an arbitrary plane: ``cos(a) m + sin(a) u`` for (P,3) medians -> (P,R,3) float32.

    The in-plane generalisation of ``generate_cone_directions`` (``src/cone.py:242-259``),
    which only produces z=0 fans.  fp64 trigonometry, fp32 result, like the reference.
    """
    m = median.double()
    m = m / m.norm(dim=-1, keepdim=True)
    u = normal_hint.double() - (normal_hint.double() * m).sum(-1, keepdim=True) * m
    u = u / u.norm(dim=-1, keepdim=True)
    a = torch.linspace(-opening_angle / 2, opening_angle / 2, n_rays, dtype=torch.float64)
    d = torch.cos(a).view(1, -1, 1) * m.unsqueeze(1) + torch.sin(a).view(1, -1, 1) * u.unsqueeze(1)
    return d.float().contiguous()


def pose_sweep(n_poses: int, n_rays: int = 128, n: int = 256, seed: int = 0,
               opening_angle: float = math.radians(60.0), radius_frac: float = 120.0 / 256.0,
               jitter_deg: float = 10.0, return_params: bool = False):
    """Config 3/4/5 poses: probes on a sphere about the volume centre looking inwards.

    Returns ``sources`` (P,3) float32 and ``directions`` (P,R,3) float32; with ``return_params`` also the (P,3)
    float32 median directions and in-plane hints the fans were built from (the inputs of ``ops.fan_directions``).
    """
    g = torch.Generator().manual_seed(seed)
    c = (n - 1) / 2.0
    v = torch.randn((n_poses, 3), generator=g, dtype=torch.float64)
    v = v / v.norm(dim=-1, keepdim=True)
    sources = c + radius_frac * n * v
    jit = torch.randn((n_poses, 3), generator=g, dtype=torch.float64)
    median = -v + math.tan(math.radians(jitter_deg)) * 0.5 * jit
    hint = torch.randn((n_poses, 3), generator=g, dtype=torch.float64)
    dirs = fan_directions(median, hint, opening_angle, n_rays)
    if return_params:
        return sources.float().contiguous(), dirs, median.float().contiguous(), hint.float().contiguous()
    return sources.float().contiguous(), dirs


def config1_pose(n: int = 256, n_rays: int = 128):
    """Config 1/2 pose: probe at the middle of the y=0 face looking along +y, 60 degree fan."""
    from .cone import generate_cone_directions
    source = torch.tensor([n / 2.0, 0.0, n / 2.0], dtype=torch.float32)
    dirs = generate_cone_directions([0.0, 1.0], math.radians(60.0), n_rays)
    return source, dirs
