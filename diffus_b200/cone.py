"""Cone / fan ray generation -- drop-in for ``generate_cone_directions`` (reference ``src/cone.py:242-259``).

The reference builds the fan on the host in float64 and casts to float32; a single fan is
R*3 numbers, so this stays a host function (bit-identical to the reference, which matters
for index-exact nearest sampling).  For batched pose sweeps the fans of all poses are
generated on the device by ``diffus_cone_directions`` (see ``ops.cone_directions``).
"""
from __future__ import annotations

import numpy as np
import torch


def generate_cone_directions(direction_mri_world, opening_angle, n_rays) -> torch.Tensor:
    """Fan of ``n_rays`` unit directions in the z=0 plane spanning ``opening_angle`` [rad].

    Same signature, dtype and values as the reference: the first two components of
    ``direction_mri_world`` are normalised to ``d``; ray ``i`` is
    ``cos(a_i) d + sin(a_i) (-d_y, d_x)`` with ``a = linspace(-angle/2, angle/2, n_rays)``
    evaluated in float64; returns ``(n_rays, 3)`` float32 with a zero third component.
    """
    d = np.asarray(direction_mri_world, dtype=np.float64).reshape(-1)[:2]
    d = d / np.linalg.norm(d)
    a = np.linspace(-opening_angle / 2, opening_angle / 2, n_rays)
    ca, sa = np.cos(a), np.sin(a)
    fan = np.zeros((n_rays, 3), dtype=np.float64)
    fan[:, 0] = ca * d[0] + sa * (-d[1])
    fan[:, 1] = ca * d[1] + sa * d[0]
    return torch.tensor(fan, dtype=torch.float32)
