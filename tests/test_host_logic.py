"""CPU: host-side logic of the drop-in layer (no kernel launches)."""
import math

import numpy as np
import pytest
import torch


def test_no_cpu_fallback_anywhere():
    import diffus_b200 as D
    from diffus_b200._lib import DiffusError
    vol = torch.rand(8, 8, 8)
    src = torch.tensor([4.0, 0.0, 4.0])
    dirs = D.generate_cone_directions([0, 1], 0.5, 4)
    ren = D.UltrasoundRenderer(16, 1e-3)
    with pytest.raises(DiffusError):
        ren.plot_beam_frame(vol, src, dirs, plot=False)
    with pytest.raises(DiffusError):
        D.render_frames(vol, src.reshape(1, 3), dirs, 16)
    with pytest.raises(DiffusError):
        D.render_mse_loss(vol, src.reshape(1, 3), dirs, torch.zeros(1, 4, 16), 16)
    with pytest.raises(DiffusError):
        D.compute_echo_traces(torch.zeros(2, 5))
    with pytest.raises(DiffusError):
        ren.simulate_rays(vol, src, dirs)
    with pytest.raises(DiffusError):
        D.custom_nearest_sampler(vol, torch.zeros(2, 5, 3))
    with pytest.raises(DiffusError):
        D.differentiable_splat(torch.zeros(4), torch.zeros(4), torch.zeros(4), torch.zeros(4), H=8, W=8)
    with pytest.raises(DiffusError):
        D.ImpedanceEstimator.compute_impedance_volume(vol, D.ImpedanceEstimator(1))


def test_reference_error_behaviour():
    import diffus_b200 as D
    ren = D.UltrasoundRenderer(16)
    assert ren.attenuation_coeff == 0.5 and ren.num_samples == 16
    vol = torch.rand(8, 8, 8)
    with pytest.raises(ValueError):                       # single ray: the reference fails to unpack (B, N)
        ren.plot_beam_frame(vol, torch.zeros(3), torch.tensor([[0.0, 1.0, 0.0]]), plot=False)
    with pytest.raises(NotImplementedError):
        ren.plot_beam_frame(vol, torch.zeros(3), torch.zeros(2, 3), artifacts=True)
    with pytest.raises(ValueError):
        D.render_frames(vol, torch.zeros(1, 3), torch.zeros(2, 3), 8, sampler="cubic")
    with pytest.raises(ValueError):
        D.compute_echo_traces(torch.zeros(5))             # (B, N) unpacking, as in the reference


def test_start_resolution_matches_reference_rules():
    from diffus_b200.renderer import _resolve_start
    assert _resolve_start(0, 100) == 0
    assert _resolve_start(-5, 100) == 0                   # clamped (reference :239-240)
    assert _resolve_start(0.25, 220) == 55                # float = fraction of num_samples (:237-238)
    assert _resolve_start(0.999, 10) == 9
    assert _resolve_start(70, 220) == 70


def test_pose_promotion_rules():
    """`source + steps * directions` dtype rules of torch (steps is float32), reference :119-124."""
    from diffus_b200.renderer import _canon_pose
    cpu = torch.device("cpu")
    s, d, pf = _canon_pose(torch.tensor([1, 2, 3]), torch.rand(4, 3), cpu)            # int64 source -> float32
    assert s.dtype == d.dtype == torch.float32 and not pf
    s, d, pf = _canon_pose(torch.rand(3, dtype=torch.float64), torch.rand(4, 3), cpu)  # product rounds to float32
    assert s.dtype == d.dtype == torch.float64 and pf
    s, d, pf = _canon_pose(torch.rand(3), torch.rand(4, 3, dtype=torch.float64), cpu)
    assert s.dtype == d.dtype == torch.float64 and not pf
    # the rule itself, checked against torch: fp64 source + fp32 directions
    src = torch.tensor([0.1, 0.2, 0.3], dtype=torch.float64)
    dirs = torch.rand(2, 3)
    steps = torch.arange(5, dtype=torch.float32).view(1, -1, 1)
    ref = src + steps * dirs.unsqueeze(1)
    mine = src + (steps * dirs.unsqueeze(1)).double()
    assert ref.dtype == torch.float64 and torch.equal(ref, mine)


def test_cone_directions_properties():
    from diffus_b200 import generate_cone_directions
    d = generate_cone_directions([3.0, 4.0, 99.0], math.radians(60), 129)
    assert d.shape == (129, 3) and d.dtype == torch.float32
    assert torch.all(d[:, 2] == 0)
    np.testing.assert_allclose(d.norm(dim=1).numpy(), 1.0, atol=1e-6)
    np.testing.assert_allclose(d[64].numpy(), [0.6, 0.8, 0.0], atol=1e-7)          # median ray
    ang = torch.atan2(d[:, 1], d[:, 0])
    np.testing.assert_allclose((ang[-1] - ang[0]).item(), math.radians(60), atol=1e-6)
    assert generate_cone_directions([1, 0], 0.3, 1).shape == (1, 3)


def test_fake_kernels_shape_inference():
    """register_fake shapes (what torch.compile / meta tracing sees) without running anything."""
    from torch._subclasses.fake_tensor import FakeTensorMode
    from diffus_b200 import ops
    with FakeTensorMode():
        vol = torch.empty(16, 16, 16)
        src = torch.empty(5, 3)
        dirs = torch.empty(5, 7, 3)
        frame, prefix = torch.ops.diffus.render_fwd(vol, None, [16, 16, 16], src, dirs, 1200, 100, 0.1, 1, False, True)
        assert frame.shape == (5, 7, 1100) and prefix.shape == (5, 7, 2, 4)
        frame, prefix = torch.ops.diffus.render_fwd(vol, None, [16, 16, 16], src, dirs, 512, 0, 0.1, 0, False, True)
        assert frame.shape == (5, 7, 512) and prefix.numel() == 0
        gv, gs, gd = torch.ops.diffus.render_bwd(frame, vol, None, [16, 16, 16], src, dirs, prefix, 512, 0, 0.1, 1,
                                                 False, True, True)
        assert gv.shape == (16, 16, 16) and gs.shape == (5, 3) and gd.shape == (5, 7, 3)
        loss, fr, gv, gs, gd = torch.ops.diffus.render_mse(vol, None, [16, 16, 16], src, dirs, frame, 512, 0, 0.1, 1,
                                                           False, False, True, False)
        assert loss.shape == (1,) and fr.numel() == 0 and gv.numel() == 0 and gd.shape == (5, 7, 3)
        assert torch.ops.diffus.echo_fwd(torch.empty(3, 40)).shape == (3, 41)
        assert torch.ops.diffus.mlp_fwd(torch.empty(1153), torch.empty(100), None, 1.0, 0.0).shape == (100,)
        assert torch.ops.diffus.mlp_bwd(torch.empty(1153), torch.empty(100), None, torch.empty(100), 1.0).shape == (1153,)


def test_impedance_estimator_state_dict_is_reference_compatible(golden_mlp):
    from diffus_b200 import ImpedanceEstimator
    from diffus_b200.impedance import pack_params
    m = ImpedanceEstimator(1)
    keys = list(m.state_dict().keys())
    assert keys == ["model.0.weight", "model.0.bias", "model.2.weight", "model.2.bias", "model.4.weight", "model.4.bias"]
    assert sum(p.numel() for p in m.parameters()) == 1153 == pack_params(m).numel()
    sd = {k: torch.tensor(golden_mlp["param_" + k.replace(".", "_")]) for k in keys}
    m.load_state_dict(sd)
    from diffus_b200._lib import DiffusError
    with pytest.raises(DiffusError):                          # no CPU forward: the MLP is a CUDA kernel
        m(torch.tensor(golden_mlp["x"]))
    flat = pack_params(m)
    assert torch.equal(flat[:32], sd["model.0.weight"].reshape(-1)) and flat[-1] == sd["model.4.bias"][0]


def test_phantoms_are_deterministic_and_well_conditioned():
    from diffus_b200.phantoms import intensity_to_impedance, layered_phantom, mri_phantom, pose_sweep
    a, b = layered_phantom(32, seed=0), layered_phantom(32, seed=0)
    assert torch.equal(a, b) and a.shape == (32, 32, 32) and a.dtype == torch.float32
    assert 1.2e6 < a.min() and a.max() < 1.9e6
    z = intensity_to_impedance(mri_phantom(32, "t1"))
    assert 1.4e6 < z.min() and z.max() < 1.8e6
    s, d = pose_sweep(7, n_rays=9, n=64, seed=1)
    assert s.shape == (7, 3) and d.shape == (7, 9, 3)
    np.testing.assert_allclose(d.norm(dim=-1).numpy(), 1.0, atol=1e-6)
    centre = torch.tensor([31.5, 31.5, 31.5])
    assert (((centre - s) / (centre - s).norm(dim=1, keepdim=True)) * d[:, 4]).sum(-1).min() > 0.9   # looking inwards
