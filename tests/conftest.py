import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


@pytest.fixture(scope="session")
def golden_echo():
    return load_golden("echo_traces.npz")


@pytest.fixture(scope="session")
def golden_frames():
    return load_golden("frames_nearest.npz")


@pytest.fixture(scope="session")
def golden_tri():
    return load_golden("frames_trilinear_grad.npz")


@pytest.fixture(scope="session")
def golden_cone():
    return load_golden("cone_directions.npz")


@pytest.fixture(scope="session")
def golden_mlp():
    return load_golden("impedance_mlp.npz")


@pytest.fixture(scope="session")
def golden_impvol():
    return load_golden("impedance_volume.npz")


@pytest.fixture(scope="session")
def golden_splat():
    return load_golden("splat.npz")


# tolerance of BASELINE.json north_star: frames <= 1e-5 relative / 1e-4 absolute; gradients <= 1e-4 relative
FRAME_RTOL, FRAME_ATOL = 1e-5, 1e-4
GRAD_RTOL = 1e-4


def assert_frame_close(got, want, what=""):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    err = np.abs(got - want)
    tol = FRAME_ATOL + FRAME_RTOL * np.abs(want)
    assert np.all(err <= tol), f"{what}: max err {err.max():.3e} (tol {tol.min():.1e}) at {np.unravel_index(err.argmax(), err.shape)}"


def assert_grad_close(got, want, what="", rtol=GRAD_RTOL):
    """Gradient tolerance: relative to the largest entry of the reference gradient (fp32 kernels vs fp64 autograd)."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    scale = max(np.abs(want).max(), 1e-30)
    err = np.abs(got - want).max()
    assert err <= rtol * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e} (rel {err / scale:.2e} > {rtol})"
