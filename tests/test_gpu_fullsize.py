"""Parity AT THE BASELINE SIZES (BASELINE.json configs 1, 2, 3, 5), through the kernels the benchmark times.

* config 1: a fixture made by the UNMODIFIED reference at full size (256^3 layered phantom, 128 rays x 512 samples,
  fp32 and fp64 runs, ~4 minutes of its dense solves; ``oracle/make_golden.py::config1_full_cases``);
* config 2: (a) the reference itself with its trilinear sampler on the same scene at 128 samples -- the deepest ray its
  autograd can differentiate in this container -- frame, d/dsource, d/ddirections, d/dvolume; (b) the full 128 x 512 step
  against the fp64 closed-form port (``oracle/port.py``, pinned to the reference by tests/test_oracle_*.py) + autograd;
* configs 3 and 5: poses drawn from the benchmark's own sweeps (256^3 MRI-shaped volume, 128 x 512; 512^3 layered
  volume, 512 x 2048, BRICK layout) against the fp64 port + autograd, through the fused one-pass / multi-pass kernel.

The volumes are regenerated from their seeds (64 MiB and 512 MiB are not committed) and checked against the fingerprint
stored with the fixture.  Tolerances: tests/conftest.py (north_star's).
"""
import numpy as np
import pytest
import torch

from conftest import assert_frame_close, assert_grad_close, load_golden, oracle_grads

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def _fingerprint(vol):
    v = vol.double().reshape(-1)
    idx = torch.arange(0, v.numel(), 104729)
    return np.array([v.sum().item(), v.square().sum().item(), (v[idx] * torch.arange(1, idx.numel() + 1)).sum().item()])


def _config1_scene(g):
    from diffus_b200.phantoms import layered_phantom
    vol = layered_phantom(256, seed=0)
    np.testing.assert_allclose(_fingerprint(vol), g["volume_fingerprint"], rtol=1e-12,
                               err_msg="layered_phantom(256, seed=0) is not the volume the reference rendered")
    return vol, torch.tensor(g["source"]), torch.tensor(g["dirs"])


@pytest.mark.parametrize("layout", [None, "brick"])
def test_config1_full_size_vs_reference_fixture(layout):
    """BASELINE config 1, the reference's own output at full size: indices bit-exact, frame within tolerance."""
    from diffus_b200 import PreparedVolume, UltrasoundRenderer, render_mse_loss
    g = load_golden("config1_full.npz")
    vol, src, dirs = _config1_scene(g)
    v = vol.to(dev())
    vv = PreparedVolume(v, layout) if layout else v
    ren = UltrasoundRenderer(int(g["S"]), float(g["alpha"]))
    x, y, z, frame = ren.plot_beam_frame(volume=vv, source=src.to(dev()), directions=dirs.to(dev()), plot=False, artifacts=False)
    assert frame.shape == (128, 512) and frame.dtype == torch.float32
    np.testing.assert_array_equal(x.cpu().numpy(), g["x"].astype(np.int64))
    np.testing.assert_array_equal(y.cpu().numpy(), g["y"].astype(np.int64))
    np.testing.assert_array_equal(z.cpu().numpy(), g["z"].astype(np.int64))
    assert_frame_close(frame.cpu().numpy(), g["frame64"], "config 1 vs reference fp64")
    # the reference's own fp32 run sits inside the same band around its fp64 run
    assert_frame_close(g["frame32"], g["frame64"], "config 1: reference fp32 vs fp64")
    # the fused one-pass kernel (the one the benchmark times) forms the same frame
    loss, fr = render_mse_loss(vv, src.to(dev()).reshape(1, 3), dirs.to(dev()), torch.zeros((1, 128, 512), device=dev()),
                               512, float(g["alpha"]), 0, sampler="nearest", return_frame=True)
    assert_frame_close(fr[0].detach().cpu().numpy(), g["frame64"], "config 1 through the fused kernel")
    np.testing.assert_allclose(loss.item(), float(np.mean(np.square(g["frame64"]))), rtol=1e-4)


@pytest.mark.parametrize("layout", [None, "brick"])
def test_config2_reduced_vs_reference_fixture(layout):
    """Config 2's scene through the REAL reference (trilinear sampler installed, fp64 autograd) at 128 samples."""
    from diffus_b200 import PreparedVolume, UltrasoundRenderer, render_mse_loss
    g = load_golden("config2_reduced.npz")
    vol, src, dirs = _config1_scene(g)
    S, alpha = int(g["S"]), float(g["alpha"])
    w = torch.tensor(g["w"], device=dev())
    # the reference-made gradients are the target; the port run in both precisions gives the fp32 noise of the same arithmetic
    from oracle import port
    w64 = torch.tensor(g["w"]).double()
    _, noise, _ = oracle_grads(lambda v_, s_, d_: (port.plot_beam_frame(v_, s_, d_, S, alpha, sampler="trilinear")[3] * w64.to(v_.dtype)).sum(),
                               vol, src, dirs)
    want_v = np.zeros(256 ** 3)
    want_v[g["grad_volume_index"]] = g["grad_volume_value"]
    want_v = want_v.reshape(256, 256, 256)

    def inputs():
        v = vol.to(dev()).requires_grad_(True)
        return v, (PreparedVolume(v, layout) if layout else v), src.to(dev()).requires_grad_(True), dirs.to(dev()).requires_grad_(True)

    # (1) forward kernel + explicit-gradient backward kernel (what autograd through plot_beam_frame runs)
    v, vv, s, d = inputs()
    _, _, _, frame = UltrasoundRenderer(S, alpha).plot_beam_frame(vv, s, d, plot=False, sampler="trilinear", return_indices=False)
    assert_frame_close(frame.detach().cpu().numpy(), g["frame64"], "config 2 (128 samples) frame")
    (frame * w).sum().backward()
    assert_grad_close(s.grad.cpu().numpy(), g["grad_source"], "d/dsource (unfused)", noise=noise[1])
    assert_grad_close(d.grad.cpu().numpy(), g["grad_dirs"], "d/ddirections (unfused)", noise=noise[2])
    assert_grad_close(v.grad.cpu().numpy(), want_v, "d/dvolume (unfused)", noise=noise[0])
    # (2) the fused one-pass kernel: with target = frame - (n/2) w the MSE's upstream gradient (2/n)(frame - target)
    # is w, so its gradients are those of sum(frame * w)
    n = frame.numel()
    target = (torch.tensor(g["frame64"], device=dev()) - 0.5 * n * w.double()).float().unsqueeze(0)
    v, vv, s, d = inputs()
    loss = render_mse_loss(vv, s.reshape(1, 3), d, target, S, alpha, 0, sampler="trilinear")
    loss.backward()
    assert_grad_close(s.grad.cpu().numpy(), g["grad_source"], "d/dsource (fused)", noise=noise[1])
    assert_grad_close(d.grad.cpu().numpy(), g["grad_dirs"], "d/ddirections (fused)", noise=noise[2])
    assert_grad_close(v.grad.cpu().numpy(), want_v, "d/dvolume (fused)", noise=noise[0])


def _port_step(vol, sources, dirs, target, S, alpha):
    """Closed-form port + autograd in float64 (and float32, for the noise of the reference's own arithmetic): frames,
    mean-squared loss against ``target``, its pose gradients, and per-pose noise of the two gradients."""
    from oracle import port
    vols = {torch.float64: vol.double(), torch.float32: vol.float()}

    def oracle(s_, d_):
        frames = torch.stack([port.plot_beam_frame(vols[s_.dtype], s_[p], d_[p], S, alpha, sampler="trilinear")[3]
                              for p in range(sources.shape[0])])
        return (frames - target.to(s_.dtype)).square().mean(), frames.detach()
    out = {}
    for dt in (torch.float64, torch.float32):
        s_, d_ = sources.to(dt).requires_grad_(True), dirs.to(dt).requires_grad_(True)
        loss, frames = oracle(s_, d_)
        out[dt] = (frames, loss.item()) + tuple(torch.autograd.grad(loss, [s_, d_]))
    f64, l64, gs, gd = out[torch.float64]
    noise_s = (out[torch.float32][2].double() - gs).abs().amax(dim=1)                 # per pose
    noise_d = (out[torch.float32][3].double() - gd).abs().amax(dim=(1, 2))
    return f64, l64, gs, gd, noise_s, noise_d


def test_config2_full_size_step_vs_port():
    """BASELINE config 2 at full size: 128 x 512 on the 256^3 layered phantom, trilinear, MSE against the frame rendered
    at source + (1.5, 0, -1): frame, loss, d/dsource, d/ddirections and d/dvolume vs the fp64 port."""
    from diffus_b200 import PreparedVolume, render_frames, render_mse_loss
    from diffus_b200.phantoms import config1_pose, layered_phantom
    from oracle import port
    vol = layered_phantom(256, seed=0)
    src, dirs = config1_pose(256, 128)
    S, alpha = 512, 1e-4
    with torch.no_grad():
        target = port.plot_beam_frame(vol.double(), src.double() + torch.tensor([1.5, 0.0, -1.0], dtype=torch.float64),
                                      dirs.double(), S, alpha, sampler="trilinear")[3]

    def oracle(v_, s_, d_):
        f_ = port.plot_beam_frame(v_, s_, d_, S, alpha, sampler="trilinear")[3]
        l_ = (f_ - target.to(f_.dtype)).square().mean()
        return l_, (f_, l_)
    (gv, gs, gd), noise, (f64, l64) = oracle_grads(oracle, vol, src, dirs)
    for layout in (None, "brick"):
        v = vol.to(dev()).requires_grad_(True)
        s, d = src.to(dev()).requires_grad_(True), dirs.to(dev()).requires_grad_(True)
        vv = PreparedVolume(v, layout) if layout else v
        loss, frame = render_mse_loss(vv, s.reshape(1, 3), d, target.float().to(dev()).unsqueeze(0), S, alpha, 0,
                                      sampler="trilinear", return_frame=True)
        loss.backward()
        assert_frame_close(frame[0].detach().cpu().numpy(), f64.detach().numpy(), f"config 2 frame ({layout})")
        np.testing.assert_allclose(loss.item(), l64.item(), rtol=1e-4)
        assert_grad_close(s.grad.cpu().numpy(), gs.numpy(), f"config 2 d/dsource ({layout})", noise=noise[1])
        assert_grad_close(d.grad.cpu().numpy(), gd.numpy(), f"config 2 d/ddirections ({layout})", noise=noise[2])
        assert_grad_close(v.grad.cpu().numpy(), gv.numpy(), f"config 2 d/dvolume ({layout})", noise=noise[0])
        # the unfused pair of kernels on the same scene
        f2 = render_frames(vv, s.detach().reshape(1, 3), d.detach(), S, alpha, sampler="trilinear")
        assert_frame_close(f2[0].detach().cpu().numpy(), f64.detach().numpy(), f"config 2 forward kernel ({layout})")


def test_config3_sweep_poses_vs_port():
    """Poses of the benchmark's own sweep (bench.py: seed 1000, 1024 poses, 256^3 MRI-shaped volume, BRICK layout)
    through the fused one-pass kernel, checked pose by pose against the fp64 port."""
    from diffus_b200 import PreparedVolume, render_frames, render_mse_loss
    from diffus_b200.phantoms import intensity_to_impedance, mri_phantom, pose_sweep
    vol = intensity_to_impedance(mri_phantom(256, "t1", seed=0))
    src, dirs = pose_sweep(1024, 128, 256, seed=1000)
    pick = [0, 1, 257, 640, 1023]
    S, alpha = 512, 1e-4
    pv = PreparedVolume(vol.to(dev()), "brick")
    shift = torch.tensor([1.5, 0.0, -1.0])
    with torch.no_grad():
        target = render_frames(pv, (src + shift).to(dev()), dirs.to(dev()), S, alpha, sampler="trilinear")
    s = src.to(dev()).requires_grad_(True)
    d = dirs.to(dev()).requires_grad_(True)
    loss, frames = render_mse_loss(pv, s, d, target, S, alpha, 0, sampler="trilinear", return_frame=True)
    loss.backward()
    # the port sees the picked poses only; the mean over all 1024 poses rescales its gradients by len(pick) / 1024
    f64, l64, gs, gd, ns, nd = _port_step(vol, src[pick], dirs[pick], target[pick].cpu(), S, alpha)
    k = len(pick) / 1024.0
    for i, p in enumerate(pick):
        assert_frame_close(frames[p].detach().cpu().numpy(), f64[i].numpy(), f"config 3 pose {p} frame")
        assert_grad_close(s.grad[p].cpu().numpy(), k * gs[i].numpy(), f"config 3 pose {p} d/dsource", noise=k * ns[i].item())
        assert_grad_close(d.grad[p].cpu().numpy(), k * gd[i].numpy(), f"config 3 pose {p} d/ddirections", noise=k * nd[i].item())
    # the targets themselves (forward kernel) against the port
    t64 = _port_step(vol, (src + shift)[pick[:2]], dirs[pick[:2]], torch.zeros_like(target[pick[:2]]).cpu(), S, alpha)[0]
    for i, p in enumerate(pick[:2]):
        assert_frame_close(target[p].cpu().numpy(), t64[i].numpy(), f"config 3 pose {p} target (forward kernel)")


def test_config5_stress_poses_vs_port():
    """Config 5's geometry at full size (512^3 layered phantom, 512 rays x 2048 samples = four 512-column passes, BRICK
    layout): frames, loss and pose gradients of four poses of its sweep vs the fp64 port."""
    from diffus_b200 import PreparedVolume, render_frames, render_mse_loss
    from diffus_b200.phantoms import layered_phantom, pose_sweep
    vol = layered_phantom(512, seed=0)
    src, dirs = pose_sweep(4, 512, 512, seed=2)
    S, alpha = 2048, 1e-4
    pv = PreparedVolume(vol.to(dev()), "brick")
    shift = torch.tensor([1.5, 0.0, -1.0])
    with torch.no_grad():
        target = render_frames(pv, (src + shift).to(dev()), dirs.to(dev()), S, alpha, sampler="trilinear")
    s = src.to(dev()).requires_grad_(True)
    d = dirs.to(dev()).requires_grad_(True)
    loss, frames = render_mse_loss(pv, s, d, target, S, alpha, 0, sampler="trilinear", return_frame=True)
    loss.backward()
    f64, l64, gs, gd, ns, nd = _port_step(vol, src, dirs, target.cpu(), S, alpha)
    np.testing.assert_allclose(loss.item(), l64, rtol=1e-4)
    for p in range(4):
        assert_frame_close(frames[p].detach().cpu().numpy(), f64[p].numpy(), f"config 5 pose {p} frame")
        assert_grad_close(s.grad[p].cpu().numpy(), gs[p].numpy(), f"config 5 pose {p} d/dsource", noise=ns[p].item())
        assert_grad_close(d.grad[p].cpu().numpy(), gd[p].numpy(), f"config 5 pose {p} d/ddirections", noise=nd[p].item())


def test_median_ties_vs_reference_fixture():
    """start > 0 with the near field outside the volume: the first kept coefficient is exactly 0 on most rays (border
    clamp), so the median over rays is a tie.  Forward vs the reference; gradient vs the out-of-place port, whose
    ``median()`` backward spreads the gradient evenly over the tied rays (torch's rule, recorded in the fixture)."""
    from diffus_b200 import UltrasoundRenderer, render_frames
    from oracle import port
    g = load_golden("median_ties.npz")
    t = torch.tensor(g["tie_rule_input"], requires_grad=True)
    t.median().backward()
    np.testing.assert_array_equal(t.grad.numpy(), g["tie_rule_grad"])         # this torch build still ties evenly
    vol, src, dirs = torch.tensor(g["volume"]), torch.tensor(g["source"]), torch.tensor(g["dirs"])
    for name in ("tie4", "tie9", "tie12"):
        start = int(g[f"{name}_start"])
        first = g[f"{name}_first_refl"]
        med = np.sort(first)[(len(first) - 1) // 2]
        assert (first == med).sum() >= (2 if name != "tie12" else 1), "the fixture is meant to hold a tie"
        ren = UltrasoundRenderer(40, 1e-3)
        x, _, _, frame = ren.plot_beam_frame(vol.to(dev()), src.to(dev()), dirs.to(dev()), plot=False, start=start)
        np.testing.assert_array_equal(x.cpu().numpy(), g[f"{name}_x"])
        assert_frame_close(frame.cpu().numpy(), g[f"{name}_frame64"], name)
        for sampler in ("nearest", "trilinear"):
            w = torch.randn((15, 40 - start), generator=torch.Generator().manual_seed(start), dtype=torch.float64)

            def oracle(v_, s_, d_):
                f_ = port.plot_beam_frame(v_, s_, d_, 40, 1e-3, start=start, sampler=sampler)[3]
                return (f_ * w.to(f_.dtype)).sum(), f_
            want, noise, f64 = oracle_grads(oracle, vol, src, dirs)
            v = vol.to(dev()).requires_grad_(True)
            s, d = src.to(dev()).requires_grad_(True), dirs.to(dev()).requires_grad_(True)
            f = render_frames(v, s.reshape(1, 3), d, 40, 1e-3, start, sampler=sampler)[0]
            assert_frame_close(f.detach().cpu().numpy(), f64.detach().numpy(), f"{name} {sampler}")
            (f * w.float().to(dev())).sum().backward()
            assert_grad_close(v.grad.cpu().numpy(), want[0].numpy(), f"{name} {sampler} d/dvolume", noise=noise[0])
            if sampler == "trilinear":
                assert_grad_close(s.grad.cpu().numpy(), want[1].numpy(), f"{name} {sampler} d/dsource", noise=noise[1])
                assert_grad_close(d.grad.cpu().numpy(), want[2].numpy(), f"{name} {sampler} d/ddirections", noise=noise[2])
