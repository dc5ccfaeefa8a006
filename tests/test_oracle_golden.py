"""CPU: the oracle (oracle/port.py) against the golden fixtures produced by the unmodified reference."""
import math

import numpy as np
import pytest
import torch

from oracle import port


def _start(g, name):
    s = float(g[f"{name}_start"])
    return s if bool(g[f"{name}_start_is_float"]) else int(s)


def test_known_answer_echoes(golden_echo):
    g = golden_echo
    for name in ("nan_lead", "total_reflection", "air_tissue_air", "doc_example"):
        Z = torch.tensor(g[f"ka_{name}_Z"])
        r = port.reflection_coeff(Z[:, :-1], Z[:, 1:])
        np.testing.assert_array_equal(np.isnan(r.numpy()), np.isnan(g[f"ka_{name}_r"]))
        np.testing.assert_allclose(np.nan_to_num(r.numpy()), np.nan_to_num(g[f"ka_{name}_r"]), rtol=1e-15)
        np.testing.assert_array_equal(port.echo_dense_solve(r).numpy(), g[f"ka_{name}_echo"])     # bit-exact: same LAPACK calls
        np.testing.assert_allclose(port.echo_closed_form(r).numpy(), g[f"ka_{name}_echo"], rtol=0, atol=1e-14)
    # the values SURVEY.md appendix B quotes from the reference
    np.testing.assert_allclose(g["ka_doc_example_echo"][0], [0, 1 / 3, 0.21212121], atol=1e-6)
    np.testing.assert_allclose(g["ka_total_reflection_echo"][0], [0, 0.0323, -0.9355, -0.9355, -0.9355, -0.9355], atol=1e-4)
    assert (g["ka_nan_lead_echo"] == 0).all()


def test_notebook_phantom(golden_echo):
    g = golden_echo
    zt = torch.tensor(g["phantom_Z"])
    r = port.reflection_coeff(zt[:, 1:], zt[:, :-1])          # the notebook's argument order
    np.testing.assert_array_equal(r.numpy(), g["phantom_r"])
    np.testing.assert_array_equal(port.echo_dense_solve(r).numpy(), g["phantom_echo"])
    np.testing.assert_allclose(port.echo_closed_form(r).numpy(), g["phantom_echo"], atol=2e-6)
    np.testing.assert_array_equal(torch.cumsum(port.surface_return_dense(r), 1).numpy(), g["phantom_cumulative"])
    np.testing.assert_allclose(port.delays_us(r.shape[1] + 1).numpy(), g["phantom_delays"], rtol=1e-6)


@pytest.mark.parametrize("name", ["rand_a", "rand_b", "rand_c"])
def test_random_coefficients(golden_echo, name):
    g = golden_echo
    r = torch.tensor(g[f"{name}_r"])
    np.testing.assert_array_equal(port.echo_dense_solve(r).numpy(), g[f"{name}_echo64"])
    np.testing.assert_allclose(port.echo_closed_form(r).numpy(), g[f"{name}_echo64"], rtol=1e-10, atol=1e-13)
    np.testing.assert_array_equal(port.echo_dense_solve(r.float()).numpy(), g[f"{name}_echo32"])
    # the reference's own float32 run sits ~1e-6 from its float64 run: the floor any fp32 kernel is judged against
    assert np.abs(g[f"{name}_echo32"] - g[f"{name}_echo64"]).max() < 2e-5 * max(1.0, np.abs(g[f"{name}_echo64"]).max())


@pytest.mark.parametrize("name", ["a0", "a7", "afrac", "b0", "b5", "c0", "d0"])
def test_plot_beam_frame_nearest(golden_frames, name):
    g = golden_frames
    vol = torch.tensor(g[f"{name}_volume"])
    src = torch.tensor(g[f"{name}_source"])
    dirs = torch.tensor(g[f"{name}_dirs"])
    S, alpha, start = int(g[f"{name}_S"]), float(g[f"{name}_alpha"]), _start(g, name)
    x, y, z, f = port.plot_beam_frame(vol.double(), src, dirs.double(), S, alpha, start=start)
    np.testing.assert_array_equal(x.numpy(), g[f"{name}_x"])
    np.testing.assert_array_equal(y.numpy(), g[f"{name}_y"])
    np.testing.assert_array_equal(z.numpy(), g[f"{name}_z"])
    np.testing.assert_allclose(f.numpy(), g[f"{name}_frame64"], rtol=0, atol=1e-13)
    # literal algorithm, float32, bit-exact against the reference's float32 run
    _, _, _, fd = port.plot_beam_frame(vol, src, dirs, S, alpha, start=start, propagation="dense")
    np.testing.assert_array_equal(fd.numpy(), g[f"{name}_frame32"])
    xs, ys, zs, imp = port.sample_nearest(vol, port.ray_points(src, dirs, S))
    np.testing.assert_array_equal(port.reflection_coeff(imp[:, :-1], imp[:, 1:]).numpy(), g[f"{name}_refl"])


@pytest.mark.parametrize("name", ["t0", "t1", "t2"])
def test_trilinear_frames_and_gradients(golden_tri, name):
    g = golden_tri
    vol = torch.tensor(g[f"{name}_volume"]).double().requires_grad_(True)
    src = torch.tensor(g[f"{name}_source"]).double().requires_grad_(True)
    dirs = torch.tensor(g[f"{name}_dirs"]).double().requires_grad_(True)
    x, y, z, f = port.plot_beam_frame(vol, src, dirs, int(g[f"{name}_S"]), float(g[f"{name}_alpha"]), sampler="trilinear")
    np.testing.assert_allclose(f.detach().numpy(), g[f"{name}_frame64"], rtol=0, atol=1e-12)
    np.testing.assert_array_equal(x.numpy(), g[f"{name}_x"])
    gv, gs, gd = torch.autograd.grad((f * torch.tensor(g[f"{name}_w"])).sum(), [vol, src, dirs])
    for got, key in ((gv, "grad_volume"), (gs, "grad_source"), (gd, "grad_dirs")):
        want = g[f"{name}_{key}"]
        assert np.abs(got.numpy() - want).max() <= 1e-9 * max(np.abs(want).max(), 1e-30), key


def test_nearest_volume_gradient(golden_tri):
    g = golden_tri
    vol = torch.tensor(g["t0_volume"]).double().requires_grad_(True)
    _, _, _, f = port.plot_beam_frame(vol, torch.tensor(g["t0_source"]), torch.tensor(g["t0_dirs"]).double(), 36, 1e-3)
    np.testing.assert_allclose(f.detach().numpy(), g["n0_frame64"], atol=1e-13)
    (gv,) = torch.autograd.grad((f * torch.tensor(g["n0_w"])).sum(), [vol])
    assert np.abs(gv.numpy() - g["n0_grad_volume"]).max() <= 1e-9 * np.abs(g["n0_grad_volume"]).max()


def test_cone_directions(golden_cone):
    from diffus_b200.cone import generate_cone_directions
    g = golden_cone
    for i in range(5):
        d, ang, n = g[f"cone{i}_d"], float(g[f"cone{i}_angle"]), int(g[f"cone{i}_n"])
        np.testing.assert_array_equal(port.generate_cone_directions(d, ang, n).numpy(), g[f"cone{i}_dirs"])
        got = generate_cone_directions(d, ang, n)          # the product's host function
        assert got.dtype == torch.float32 and got.shape == (n, 3)
        np.testing.assert_array_equal(got.numpy(), g[f"cone{i}_dirs"])


def test_mlp(golden_mlp):
    g = golden_mlp
    p = [torch.tensor(g[k]) for k in ("param_model_0_weight", "param_model_0_bias", "param_model_2_weight",
                                      "param_model_2_bias", "param_model_4_weight", "param_model_4_bias")]
    y = port.mlp_forward(torch.tensor(g["x"]), *p)
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=1e-5, atol=1e-6)
    y64 = port.mlp_forward(torch.tensor(g["x"]).double(), *[q.double() for q in p])
    np.testing.assert_allclose(y64.numpy(), g["y64"], rtol=1e-12)


def test_splat(golden_splat):
    g = golden_splat
    x, y, z, val = (torch.tensor(g[k]) for k in ("x", "y", "z", "val"))
    for sigma in (0.5, 1.0):
        img = port.splat(x, y, z, val, H=64, W=64, sigma=sigma)
        np.testing.assert_allclose(img.numpy(), g[f"img_sigma{sigma}"], rtol=1e-6, atol=1e-7)


def test_closed_form_equals_dense_on_layered_medium():
    """The identity the CUDA kernels rest on, at a size the dense form still reaches, fp64, incl. gradients."""
    g = torch.Generator().manual_seed(0)
    layers = 1.4e6 + 0.3e6 * torch.rand((4, 8), generator=g, dtype=torch.float64)
    Z = layers.repeat_interleave(12, dim=1) * (1 + 0.005 * torch.randn((4, 96), generator=g, dtype=torch.float64))
    r = port.reflection_coeff(Z[:, :-1], Z[:, 1:]).requires_grad_(True)
    a, b = port.echo_dense_solve(r), port.echo_closed_form(r)
    np.testing.assert_allclose(a.detach().numpy(), b.detach().numpy(), rtol=0, atol=1e-13)
    w = torch.randn(a.shape, generator=g, dtype=torch.float64)
    (ga,) = torch.autograd.grad((a * w).sum(), r, retain_graph=True)
    (gb,) = torch.autograd.grad((b * w).sum(), r)
    np.testing.assert_allclose(ga.numpy(), gb.numpy(), rtol=0, atol=1e-11)


def test_preprocessing_and_impedance_volume(golden_impvol):
    g = golden_impvol
    vol = torch.tensor(g["volume"])
    mask = port.create_brain_mask(vol, 50)
    np.testing.assert_array_equal(mask.numpy(), g["mask"])
    np.testing.assert_allclose(port.zscore_normalize(vol, mask).numpy(), g["vol_norm"], rtol=1e-6, atol=1e-7)
    params = [torch.tensor(g[k]) for k in ("param_model_0_weight", "param_model_0_bias", "param_model_2_weight",
                                           "param_model_2_bias", "param_model_4_weight", "param_model_4_bias")]
    np.testing.assert_allclose(port.compute_impedance_volume(vol, params).numpy(), g["Z"], rtol=1e-5, atol=1.0)


def test_reference_known_answers_for_the_sign_convention():
    """The only numbers the reference's own notebooks pin for this path.

    notebooks/[DEMO] Intro to the theory behind propagation.ipynb cell 12: propagate_boundary(g=10, d=10, Z1=1, Z2=2)
    prints r=0.3333, tLR=1.3333, tRL=0.6667 and returns (16.6667, 10.0): r = (Z2-Z1)/(Z1+Z2), tLR = 1+r, tRL = 1-r.
    """
    r = port.reflection_coeff(torch.tensor([1.0]), torch.tensor([2.0]))
    np.testing.assert_allclose([r.item(), (1 + r).item(), (1 - r).item()], [0.3333, 1.3333, 0.6667], atol=5e-5)
    g = d = torch.tensor([10.0])
    new_g, new_d = (1 + r) * g + r * d, (1 - r) * d + r * g
    np.testing.assert_allclose([new_g.item(), new_d.item()], [16.6667, 10.0], atol=5e-5)
    # the same convention is what the layered system of src/renderer.py:380-405 encodes: one interface, d0 = r
    np.testing.assert_allclose(port.surface_return_dense(r.reshape(1, 1))[0].numpy(), [0.0, r.item()], atol=1e-7)
    # forward_physics.md:52-88 works Z=(1,2,1.5) with the PHYSICAL convention and does not pin the code;
    # the code's own answer for it (measured on the reference) is echo [0, 0.3333, 0.2121]
    Z = torch.tensor([[1.0, 2.0, 1.5]], dtype=torch.float64)
    e = port.echo_closed_form(port.reflection_coeff(Z[:, :-1], Z[:, 1:]))
    np.testing.assert_allclose(e[0].numpy(), [0.0, 1 / 3, 0.21212121], atol=1e-7)


def test_notebook_brain_phantom_2d():
    """generate_brain_phantom_2d (Modeling Choices cell 5): air / bone interfaces, |r| up to 0.9995."""
    from conftest import load_golden
    g = load_golden("echo_brain_phantom2d.npz")
    ph = torch.tensor(g["Z"])
    r = port.reflection_coeff(ph[:, 1:], ph[:, :-1])
    np.testing.assert_array_equal(r.numpy(), g["r"])
    np.testing.assert_array_equal(port.echo_dense_solve(r).numpy(), g["echo32"])
    np.testing.assert_allclose(port.echo_closed_form(r.double()).numpy(), g["echo64"], atol=1e-13)


# ----------------------------------------------------------------------------------------
# BASELINE sizes: the closed-form port against the reference's own full-size outputs
# ----------------------------------------------------------------------------------------
def _fingerprint(vol):
    v = vol.double().reshape(-1)
    idx = torch.arange(0, v.numel(), 104729)
    return np.array([v.sum().item(), v.square().sum().item(), (v[idx] * torch.arange(1, idx.numel() + 1)).sum().item()])


@pytest.fixture(scope="module")
def layered256():
    from diffus_b200.phantoms import layered_phantom
    return layered_phantom(256, seed=0)


def test_config1_full_size_port_vs_reference(layered256):
    """Config 1 at full size (128 x 512 on the 256^3 phantom): the O(S) closed form equals the reference's 512 dense
    solves per ray to 1e-13 in fp64, indices bit-exact; the reference's fp32 run is 1e-5 of peak away from both."""
    from conftest import load_golden
    g = load_golden("config1_full.npz")
    np.testing.assert_allclose(_fingerprint(layered256), g["volume_fingerprint"], rtol=1e-12)
    src, dirs = torch.tensor(g["source"]), torch.tensor(g["dirs"])
    x, y, z, f = port.plot_beam_frame(layered256.double(), src, dirs.double(), int(g["S"]), float(g["alpha"]))
    np.testing.assert_array_equal(x.numpy(), g["x"].astype(np.int64))
    np.testing.assert_array_equal(y.numpy(), g["y"].astype(np.int64))
    np.testing.assert_array_equal(z.numpy(), g["z"].astype(np.int64))
    peak = np.abs(g["frame64"]).max()
    assert 0.05 < peak < 0.1                      # frame range [-0.081, 0.067]
    np.testing.assert_allclose(f.numpy(), g["frame64"], rtol=0, atol=1e-12 * peak + 1e-13)
    assert np.abs(g["frame32"] - g["frame64"]).max() < 5e-5 * peak


def test_config2_reduced_port_vs_reference(layered256):
    """Config 2's scene, 128 samples, through the reference with its trilinear sampler: the port's frame and all three
    gradients (autograd through the closed form vs autograd through the dense solves) in fp64."""
    from conftest import load_golden
    g = load_golden("config2_reduced.npz")
    np.testing.assert_allclose(_fingerprint(layered256), g["volume_fingerprint"], rtol=1e-12)
    v64 = layered256.double().requires_grad_(True)
    s64 = torch.tensor(g["source"]).double().requires_grad_(True)
    d64 = torch.tensor(g["dirs"]).double().requires_grad_(True)
    f = port.plot_beam_frame(v64, s64, d64, int(g["S"]), float(g["alpha"]), sampler="trilinear")[3]
    np.testing.assert_allclose(f.detach().numpy(), g["frame64"], rtol=0, atol=1e-12)
    gv, gs, gd = torch.autograd.grad((f * torch.tensor(g["w"]).double()).sum(), [v64, s64, d64])
    np.testing.assert_allclose(gs.numpy(), g["grad_source"], rtol=1e-8, atol=1e-10 * np.abs(g["grad_source"]).max())
    np.testing.assert_allclose(gd.numpy(), g["grad_dirs"], rtol=1e-8, atol=1e-10 * np.abs(g["grad_dirs"]).max())
    want = np.zeros(256 ** 3)
    want[g["grad_volume_index"]] = g["grad_volume_value"]
    np.testing.assert_allclose(gv.numpy().reshape(-1), want, rtol=2e-7, atol=2e-7 * np.abs(want).max())   # stored as float32


def test_median_ties_port_vs_reference():
    """start > 0 with a tie at the median: forward vs the reference; torch's even split of the gradient over the ties."""
    from conftest import load_golden
    g = load_golden("median_ties.npz")
    vol, src, dirs = torch.tensor(g["volume"]), torch.tensor(g["source"]), torch.tensor(g["dirs"])
    for name in ("tie4", "tie9", "tie12"):
        start = int(g[f"{name}_start"])
        x, _, _, f = port.plot_beam_frame(vol.double(), src, dirs.double(), 40, 1e-3, start=start)
        np.testing.assert_array_equal(x.numpy(), g[f"{name}_x"])
        np.testing.assert_allclose(f.numpy(), g[f"{name}_frame64"], rtol=0, atol=1e-13)
    t = torch.tensor(g["tie_rule_input"], requires_grad=True)
    t.median().backward()
    np.testing.assert_array_equal(t.grad.numpy(), g["tie_rule_grad"])


# ----------------------------------------------------------------------------------------
# training-loop pieces (rows f3 / f1 epilogue)
# ----------------------------------------------------------------------------------------
def test_training_loop_port_vs_reference_fixture():
    from conftest import load_golden
    g = load_golden("training_loop.npz")
    # rotate_around_apex: the reference function's own output
    xr, zr = port.rotate_around_apex(torch.tensor(g["rot_x"]), torch.tensor(g["rot_z"]), torch.tensor(g["rot_apex"]).float(),
                                     list(g["rot_median"]))
    np.testing.assert_allclose(xr.numpy(), g["rot_x_out"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(zr.numpy(), g["rot_z_out"], rtol=0, atol=1e-4)
    # masked MSE + edge loss: the CPU notebook's cell, executed
    a = torch.tensor(g["mse_edge_synth"], requires_grad=True)
    loss = port.masked_mse_edge_loss(a, torch.tensor(g["mse_edge_real"]), torch.tensor(g["mse_edge_mask"]))
    np.testing.assert_allclose(loss.item(), float(g["mse_edge_loss"]), rtol=1e-14)
    (ga,) = torch.autograd.grad(loss, a)
    np.testing.assert_allclose(ga.numpy(), g["mse_edge_grad"], rtol=1e-12, atol=1e-18)
    # process_rf_to_bmode: the notebook's cell with scipy.signal.hilbert vs the numpy-FFT restatement
    for name in ("rf_even", "rf_odd"):
        np.testing.assert_allclose(port.process_rf_to_bmode(torch.tensor(g[name])), g[name + "_bmode"], rtol=1e-5, atol=1e-7)
    # SSIM restatement: sanity properties (piq itself is not available: parity unpinned, said so in the fixture)
    assert "parity unpinned" in str(g["ssim_made_by"])
    y = torch.tensor(g["ssim_real"])
    assert abs(port.ssim_piq(y, y).item() - 1.0) < 1e-12
    assert port.ssim_piq(y, 1 - y).item() < 0.2


def _zigzag_mlp():
    """Weights whose network has > 1000 linear pieces: layer-1 units switch at 0, 1, ..., 31 and every layer-2 unit zigzags
    through zero once per interval (the CUDA path falls back to the layered evaluation above 256 regions)."""
    import torch
    w1 = torch.ones(32, 1)
    b1 = -torch.arange(32, dtype=torch.float32)
    w2 = torch.zeros(32, 32)
    for j in range(32):
        w2[j, 0] = 1.0
        w2[j, 1:] = 2.0 * (-1.0) ** torch.arange(1, 32)          # slopes +1, -1, +1, ... from x = 0 on
    b2 = -(0.2 + 0.6 * torch.arange(32, dtype=torch.float32) / 32.0)   # a different crossing height per unit
    w3 = torch.linspace(-1.0, 1.0, 32).reshape(1, 32)
    b3 = torch.tensor([0.25])
    return [w1, b1, w2, b2, w3, b3]


@pytest.mark.parametrize("seed", [0, 1, 2, "zigzag"])
def test_mlp_piecewise_linear_restatement_matches_layered_oracle(seed):
    """The piecewise-linear form of the scalar-input MLP (what DIFFUS_MLP_PATH_PIECEWISE evaluates) == the layered network
    (src/impedance.py:10-17), values and weight gradients, in float64."""
    import torch
    from oracle import port
    if seed == "zigzag":
        prm = [p.double() for p in _zigzag_mlp()]
    else:
        torch.manual_seed(seed)
        m = torch.nn.Sequential(torch.nn.Linear(1, 32), torch.nn.ReLU(), torch.nn.Linear(32, 32), torch.nn.ReLU(), torch.nn.Linear(32, 1))
        prm = [p.detach().double() for p in m.parameters()]
    gen = torch.Generator().manual_seed(7)
    x = (torch.rand(20000, generator=gen, dtype=torch.float64) * 40.0 - 4.0) if seed == "zigzag" else \
        torch.randn(20000, generator=gen, dtype=torch.float64) * 3.0
    g = torch.randn(20000, generator=gen, dtype=torch.float64)
    bps, P, Q, _ = port.mlp_piecewise_table(*[p.numpy() for p in prm])
    assert (len(bps) > 1000) if seed == "zigzag" else (32 <= len(bps) <= 1088)
    r = np.searchsorted(bps, x.numpy(), side="right")
    want = port.mlp_forward(x.reshape(-1, 1), *prm).reshape(-1)
    np.testing.assert_allclose(P[r] * x.numpy() + Q[r], want.numpy(), rtol=0, atol=1e-12 * max(1.0, float(want.abs().max())))
    leaves = [p.clone().requires_grad_(True) for p in prm]
    (port.mlp_forward(x.reshape(-1, 1), *leaves).reshape(-1) * g).sum().backward()
    got = port.mlp_piecewise_grads(x.numpy(), g.numpy(), *[p.numpy() for p in prm])
    for a, b in zip(got, leaves):
        np.testing.assert_allclose(a, b.grad.numpy(), rtol=1e-9, atol=1e-9 * float(b.grad.abs().max()))
