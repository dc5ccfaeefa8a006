"""CPU, build container only: the oracle against the LIVE reference (skipped where /root/reference is absent)."""
import math

import numpy as np
import pytest
import torch

from oracle import port
from oracle import reference_loader as RL

pytestmark = pytest.mark.skipif(not RL.available(), reason="reference tree not mounted (expected on the GPU box)")


@pytest.fixture(scope="module")
def ref():
    return RL.load()


def test_echo_traces_live(ref):
    g = torch.Generator().manual_seed(11)
    r = 0.01 * torch.randn((6, 60), generator=g, dtype=torch.float64)
    with RL.quiet():
        e, d = ref.renderer.compute_echo_traces(r)
    np.testing.assert_array_equal(port.echo_dense_solve(r).numpy(), e.numpy())
    np.testing.assert_allclose(port.echo_closed_form(r).numpy(), e.numpy(), atol=1e-14)


@pytest.mark.parametrize("start", [0, 9, 0.3])
def test_plot_beam_frame_live(ref, start):
    from diffus_b200.phantoms import layered_phantom
    vol = layered_phantom(40, seed=2).double()
    src = torch.tensor([20.0, 0.0, 20.0], dtype=torch.float64)
    dirs = ref.cone.generate_cone_directions([0.1, 1.0], math.radians(55), 10).double()
    with RL.quiet():
        x, y, z, f = ref.renderer.UltrasoundRenderer(56, 1e-3).plot_beam_frame(
            volume=vol.clone(), source=src, directions=dirs, plot=False, artifacts=False, start=start)
    xo, yo, zo, fo = port.plot_beam_frame(vol, src, dirs, 56, 1e-3, start=start)
    assert torch.equal(x, xo) and torch.equal(y, yo) and torch.equal(z, zo)
    np.testing.assert_allclose(fo.numpy(), f.numpy(), atol=1e-13)


def test_trilinear_gradients_live(ref):
    from diffus_b200.phantoms import layered_phantom, pose_sweep
    vol = layered_phantom(24, seed=4).double().requires_grad_(True)
    sources, dirs = pose_sweep(1, n_rays=6, n=24, seed=3)
    s = sources[0].double().requires_grad_(True)
    d = dirs[0].double().requires_grad_(True)
    with RL.trilinear_sampler_installed(ref), RL.quiet():
        _, _, _, f = ref.renderer.UltrasoundRenderer(40, 1e-3).plot_beam_frame(volume=vol, source=s, directions=d, plot=False)
    _, _, _, fo = port.plot_beam_frame(vol, s, d, 40, 1e-3, sampler="trilinear")
    np.testing.assert_allclose(fo.detach().numpy(), f.detach().numpy(), atol=1e-12)
    w = torch.randn(f.shape, generator=torch.Generator().manual_seed(0), dtype=torch.float64)
    ga = torch.autograd.grad((f * w).sum(), [vol, s, d])
    gb = torch.autograd.grad((fo * w).sum(), [vol, s, d])
    for a, b in zip(ga, gb):
        assert (a - b).abs().max() <= 1e-9 * a.abs().max()


def test_mlp_and_splat_live(ref):
    torch.manual_seed(1)
    model = ref.impedance.ImpedanceEstimator(1)
    x = torch.randn(50, 1)
    p = [q.detach() for q in model.parameters()]
    np.testing.assert_allclose(port.mlp_forward(x, *p).numpy(), model(x).detach().numpy(), rtol=1e-5, atol=1e-6)
    k = torch.arange(30).float()
    th = torch.linspace(-0.3, 0.3, 8)
    xs = (20 + k[None] * torch.sin(th)[:, None]).round().long()
    ys = (1 + k[None] * torch.cos(th)[:, None]).round().long()
    zs = torch.full_like(xs, 5)
    val = torch.randn(8, 30)
    with RL.quiet():
        img = ref.renderer.differentiable_splat(xs, ys, zs, val, H=48, W=48, sigma=1.0)
    np.testing.assert_allclose(port.splat(xs, ys, zs, val, H=48, W=48, sigma=1.0).numpy(), img.numpy(), rtol=1e-6, atol=1e-7)


def test_rotate_around_apex_live(ref):
    from oracle.port import rotate_around_apex          # the product's version is a CUDA kernel (tests/test_gpu_training.py)
    g = torch.Generator().manual_seed(2)
    x = torch.rand(7, 11, generator=g) * 200
    z = torch.rand(7, 11, generator=g) * 200
    want = ref.renderer.rotate_around_apex(x.reshape(-1), z.reshape(-1), (120.0, 4.0), (0.3, -0.9))
    got = rotate_around_apex(x.reshape(-1), z.reshape(-1), (120.0, 4.0), (0.3, -0.9))
    np.testing.assert_allclose(got[0].numpy(), want[0].numpy(), rtol=1e-6, atol=1e-4)
    np.testing.assert_allclose(got[1].numpy(), want[1].numpy(), rtol=1e-6, atol=1e-4)


def test_gaussian_pulse_live(ref):
    from diffus_b200 import gaussian_pulse
    for length, sigma in ((10, 1), (20, 4), (7, 2.5)):
        np.testing.assert_array_equal(gaussian_pulse(length, sigma), ref.renderer.gaussian_pulse(length, sigma))
