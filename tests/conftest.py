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


# Tolerances of BASELINE.json north_star:
#   frames    <= 1e-5 relative / 1e-4 absolute ON NORMALISED INTENSITY: the absolute term is scaled by the peak of the
#             reference frame (a raw B-mode line peaks at ~0.08 for tissue, at 1e3 for resonating air gaps);
#   gradients <= 1e-4 relative vs (fp64) torch autograd, ELEMENT-WISE for every entry above GRAD_FLOOR of the largest
#             one; smaller entries (sums that cancel) are held to the same absolute error as an entry at the floor.
FRAME_RTOL, FRAME_ATOL = 1e-5, 1e-4
GRAD_RTOL, GRAD_FLOOR = 1e-4, 1e-3

# DIFFUS_TOL_REPORT=<file>: every check appends (test, what, achieved error / tolerance) -- how the margins in
# profiles/r2_parity_margins.md were obtained.  DIFFUS_TOL_CALIBRATE=1 additionally records instead of failing.
_REPORT = os.environ.get("DIFFUS_TOL_REPORT")
_CALIBRATE = os.environ.get("DIFFUS_TOL_CALIBRATE") == "1"


def _record(kind, what, ratio, detail):
    if _REPORT:
        import json
        test = os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0]
        with open(_REPORT, "a") as f:
            f.write(json.dumps({"test": test, "kind": kind, "what": what, "err_over_tol": ratio, **detail}) + "\n")


def assert_frame_close(got, want, what="", atol=FRAME_ATOL, rtol=FRAME_RTOL):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    if got.size == 0:
        return
    peak = float(np.abs(want).max())
    err = np.abs(got - want)
    tol = atol * peak + rtol * np.abs(want) + 1e-30
    ratio = float((err / tol).max())
    _record("frame", what, ratio, {"max_err": float(err.max()), "peak": peak})
    if _CALIBRATE:
        return
    assert ratio <= 1.0, (f"{what}: max err {err.max():.3e} at {np.unravel_index((err / tol).argmax(), err.shape)}, "
                          f"{ratio:.2f} x the tolerance ({atol:g} x peak {peak:.3e} + {rtol:g} |ref|)")


def assert_grad_close(got, want, what="", rtol=GRAD_RTOL, floor=GRAD_FLOOR):
    """fp32 kernels vs fp64 autograd: |got - want| <= rtol |want| for every entry with |want| > floor * max|want|,
    and <= rtol * floor * max|want| for the entries below that."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    if got.size == 0:
        return
    scale = max(float(np.abs(want).max()), 1e-300)
    err = np.abs(got - want)
    tol = rtol * np.maximum(np.abs(want), floor * scale)
    ratio = float((err / tol).max())
    _record("grad", what, ratio, {"max_err_over_scale": float(err.max() / scale), "scale": scale})
    if _CALIBRATE:
        return
    k = np.unravel_index((err / tol).argmax(), err.shape)
    assert ratio <= 1.0, (f"{what}: entry {k}: got {got[k]:.6e} want {want[k]:.6e} (|ref| max {scale:.3e}); "
                          f"{ratio:.2f} x the tolerance (rtol {rtol:g}, floor {floor:g})")
