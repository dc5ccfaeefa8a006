"""Import the UNMODIFIED reference (gduguey/DiffUS) for oracle checking.  TEST INFRASTRUCTURE.

The reference tree lives at ``/root/reference`` in the build container only.  For the CPU-baseline
legs of ``bench.py`` (which must run the reference itself on the GPU box's host cores),
``__graft_entry__.build()`` packs its four hot-path modules, byte for byte, into the build output
``oracle/_ref/reference_src.zip`` (git-ignored: never part of this repository's tree or history; it travels to
the GPU box like a built ``.so``).  :func:`load` looks at ``/root/reference`` first and unpacks the archive into a
temporary directory second; everything that calls it must still tolerate ``None``.

Three harness-level shims, none of which alters arithmetic:

1. the plotting / IO modules the reference imports at module top but which are not
   installed here (``src/renderer.py:7,13``, ``src/cone.py:3-4``, ``src/utils.py:1-4``)
   are replaced by ``MagicMock`` modules;
2. ``custom_nearest_sampler`` is re-bound with ``visualize=False`` -- the debug
   visualiser is hard-wired on (``src/renderer.py:741``) and calls ``Z.cpu().numpy()``
   on grad-requiring tensors (``src/renderer.py:783``); the name is resolved at call time
   inside ``trace_ray`` (``src/renderer.py:178``);
3. the unconditional ``print`` calls are swallowed by :func:`quiet`.
"""
from __future__ import annotations

import contextlib
import functools
import io
import os
import sys
import types
from unittest.mock import MagicMock

STAGED_ARCHIVE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference_src.zip")
_HOT_PATH_MODULES = ("renderer.py", "cone.py", "impedance.py", "utils.py")
_extracted: str | None = None


def _find_root() -> str:
    """``/root/reference`` (build container), else the staged archive unpacked into a temporary directory."""
    global _extracted
    env = os.environ.get("DIFFUS_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile(os.path.join("/root/reference", "src", "renderer.py")):
        return "/root/reference"
    if os.path.isfile(STAGED_ARCHIVE):
        import atexit
        import shutil
        import tempfile
        import zipfile
        _extracted = tempfile.mkdtemp(prefix="diffus_reference_")
        atexit.register(shutil.rmtree, _extracted, ignore_errors=True)
        with zipfile.ZipFile(STAGED_ARCHIVE) as z:
            z.extractall(_extracted)
        return _extracted
    return "/root/reference"


def stage(source_root: str = "/root/reference") -> str | None:
    """Pack the reference's hot-path modules, byte for byte, into ``oracle/_ref/reference_src.zip`` so that the CPU
    baseline can run the reference itself on a machine without ``/root/reference``.  The archive is a build output
    (git-ignored, like a compiled ``.so``): no reference source file is ever placed in this repository's tree."""
    import zipfile
    src = os.path.join(source_root, "src")
    if not os.path.isfile(os.path.join(src, "renderer.py")):
        return None
    os.makedirs(os.path.dirname(STAGED_ARCHIVE), exist_ok=True)
    with zipfile.ZipFile(STAGED_ARCHIVE, "w", zipfile.ZIP_DEFLATED) as z:
        for name in _HOT_PATH_MODULES:
            z.write(os.path.join(src, name), os.path.join("src", name))
        lic = os.path.join(source_root, "LICENSE")
        if os.path.isfile(lic):
            z.write(lic, "LICENSE")
    return STAGED_ARCHIVE


def is_staged_copy() -> bool:
    return _extracted is not None


REFERENCE_ROOT = _find_root()

_STUBS = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.widgets", "matplotlib.animation",
    "nibabel", "plotly", "plotly.graph_objects", "plotly.io", "torchio", "jax", "jax.numpy",
]

_cache: types.SimpleNamespace | None = None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "renderer.py"))


def load():
    """Return a namespace ``(renderer, cone, impedance, utils)`` of reference modules, or None."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        return None
    for name in _STUBS:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = MagicMock(name=name)
                m.__path__ = []
                sys.modules[name] = m
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    with quiet():
        renderer = importlib.import_module("src.renderer")
        cone = importlib.import_module("src.cone")
        impedance = importlib.import_module("src.impedance")
        utils = importlib.import_module("src.utils")
    if not isinstance(renderer.custom_nearest_sampler, functools.partial):
        renderer._orig_custom_nearest_sampler = renderer.custom_nearest_sampler
        renderer.custom_nearest_sampler = functools.partial(
            renderer.custom_nearest_sampler, visualize=False)
    _cache = types.SimpleNamespace(renderer=renderer, cone=cone, impedance=impedance, utils=utils)
    return _cache


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


@contextlib.contextmanager
def trilinear_sampler_installed(ref):
    """Swap the reference's sampler for its own notebook-era trilinear variant.

    The variant is the ``F.grid_sample(mode='bilinear', padding_mode='border',
    align_corners=True)`` form with grid x<-p2, y<-p1, z<-p0 found in
    ``notebooks/[DEPR] fxiafixing_voxel_plot.ipynb`` cell 29 and
    ``notebooks/[DEMO] Renderer Alternatives.ipynb`` cell 6; HEAD keeps the same grid
    construction for its ``mode='nearest'`` branch (``src/renderer.py:802-815``).  It is the
    only form in which the reference has gradients w.r.t. the probe pose.
    """
    import torch
    import torch.nn.functional as F

    def sampler(Z, points, visualize=False, sampler="prop", start=0):
        D, H, W = Z.shape
        pts = points.to(Z.dtype)
        B, S, _ = pts.shape
        grid = torch.stack([2 * pts[..., 2] / (W - 1) - 1,
                            2 * pts[..., 1] / (H - 1) - 1,
                            2 * pts[..., 0] / (D - 1) - 1], -1).view(1, B, S, 1, 3)
        v = F.grid_sample(Z[None, None], grid, mode="bilinear", padding_mode="border",
                          align_corners=True).view(B, S)
        x = torch.clamp(pts[..., 0].round().long(), 0, D - 1)
        y = torch.clamp(pts[..., 1].round().long(), 0, H - 1)
        z = torch.clamp(pts[..., 2].round().long(), 0, W - 1)
        return x, y, z, v

    saved = ref.renderer.custom_nearest_sampler
    ref.renderer.custom_nearest_sampler = sampler
    try:
        yield
    finally:
        ref.renderer.custom_nearest_sampler = saved
