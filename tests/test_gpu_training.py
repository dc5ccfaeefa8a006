"""GPU: the training-loop pieces around the path (SURVEY.md 8 rows f3 / f1 epilogue) against fixtures made from the
reference (its ``rotate_around_apex``; the notebooks' loss and B-mode cells, executed; ``torch.optim.Adam``) and against
``oracle/port.py`` (the SSIM restatement of piq -- piq is not installed anywhere: that one is parity-unpinned)."""
import numpy as np
import pytest
import torch

from conftest import oracle_grads, assert_frame_close, assert_grad_close, load_golden

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def test_rotate_around_apex_matches_reference():
    from diffus_b200 import rotate_around_apex
    g = load_golden("training_loop.npz")
    xr, zr = rotate_around_apex(torch.tensor(g["rot_x"], device=dev()), torch.tensor(g["rot_z"], device=dev()),
                                torch.tensor(g["rot_apex"]).float(), list(g["rot_median"]))
    # the reference rotates with a 2 x 2 matrix product (sgemm rounding, possibly fused): agreement to 1 ulp of the coordinates
    np.testing.assert_allclose(xr.cpu().numpy(), g["rot_x_out"], rtol=0, atol=5e-5)
    np.testing.assert_allclose(zr.cpu().numpy(), g["rot_z_out"], rtol=0, atol=5e-5)
    x2 = torch.tensor(g["rot_x"], device=dev()).reshape(12, 40)                  # (rays, samples) arrays, as the notebooks pass
    xr2, _ = rotate_around_apex(x2, torch.tensor(g["rot_z"], device=dev()).reshape(12, 40), g["rot_apex"], list(g["rot_median"]))
    assert xr2.shape == (12, 40) and torch.equal(xr2.reshape(-1), xr)


def test_masked_mse_edge_loss_matches_notebook_cell():
    from diffus_b200.losses import masked_mse_edge_loss
    g = load_golden("training_loop.npz")
    a = torch.tensor(g["mse_edge_synth"], device=dev(), dtype=torch.float32, requires_grad=True)
    loss = masked_mse_edge_loss(a, torch.tensor(g["mse_edge_real"], device=dev(), dtype=torch.float32),
                                torch.tensor(g["mse_edge_mask"], device=dev()))
    np.testing.assert_allclose(loss.item(), float(g["mse_edge_loss"]), rtol=1e-5)
    (3.0 * loss).backward()
    assert_grad_close(a.grad.cpu().numpy(), 3.0 * g["mse_edge_grad"], "d (masked mse + edge) / d synth")


@pytest.mark.parametrize("normalize", [True, False])
def test_ssim_loss_matches_port(normalize):
    from diffus_b200.losses import ssim_loss
    from oracle import port
    g = load_golden("training_loop.npz")
    tag = "ssim_norm" if normalize else "ssim_raw"
    s = torch.tensor(g["ssim_synth"], device=dev(), dtype=torch.float32, requires_grad=True)
    y = torch.tensor(g["ssim_real"], device=dev(), dtype=torch.float32)
    loss = ssim_loss(s, y, normalize=normalize)
    np.testing.assert_allclose(loss.item(), float(g[tag + "_loss"]), rtol=2e-5)
    (2.0 * loss).backward()
    # SSIM's local variances are differences of window means (E[x^2] - mu^2): the reference's own float32 evaluation is that
    # far from its float64 one, so the element-wise 1e-4 is widened by the measured float32 drift of the port (oracle_grads)
    y_cpu = torch.tensor(g["ssim_real"])
    _, noise, _ = oracle_grads(lambda s_: 2.0 * port.ssim_loss(s_, y_cpu.to(s_.dtype), normalize=normalize), torch.tensor(g["ssim_synth"]))
    assert_grad_close(s.grad.cpu().numpy(), 2.0 * g[tag + "_grad"], f"d (1 - ssim) / d synth ({tag})", noise=noise[0])
    # a 256 x 256 image (the notebooks' size) against the port run on the host, other window parameters too
    gen = torch.Generator().manual_seed(5)
    real = torch.rand((256, 256), generator=gen)
    synth = (real * 0.7 + 0.2 * torch.rand((256, 256), generator=gen)).requires_grad_(True)
    fn = lambda s_: (lambda l_: (l_, l_))(port.ssim_loss(s_, real.to(s_.dtype), normalize=normalize, kernel_size=7, kernel_sigma=1.0))
    (gw,), noise, want = oracle_grads(fn, synth)
    sd = synth.detach().to(dev()).requires_grad_(True)
    got = ssim_loss(sd, real.to(dev()), normalize=normalize, kernel_size=7, kernel_sigma=1.0)
    np.testing.assert_allclose(got.item(), want.item(), rtol=2e-5)
    got.backward()
    assert_grad_close(sd.grad.cpu().numpy(), gw.numpy(), "d (1 - ssim) / d synth, 256 x 256", noise=noise[0])


def test_log_compression():
    from diffus_b200.losses import log_compress, process_rf_to_bmode
    from oracle import port
    g = load_golden("training_loop.npz")
    for name in ("rf_even", "rf_odd"):                                  # the notebook cell's own output (scipy.signal.hilbert)
        got = process_rf_to_bmode(torch.tensor(g[name], device=dev()))
        np.testing.assert_allclose(got.cpu().numpy(), g[name + "_bmode"], rtol=2e-5, atol=2e-6)
    gen = torch.Generator().manual_seed(2)
    rf = 0.05 * torch.randn((128, 512), generator=gen)                  # a full frame
    np.testing.assert_allclose(process_rf_to_bmode(rf.to(dev())).cpu().numpy(), port.process_rf_to_bmode(rf), rtol=5e-5, atol=5e-6)
    img = torch.randn((64, 48), generator=gen)
    img[3, 4] = img[10, 11] = 7.5                                       # a tie at the maximum
    x64 = img.double().requires_grad_(True)
    want = port.log_compress(x64)
    w = torch.randn(want.shape, generator=gen, dtype=torch.float64)
    (gw,) = torch.autograd.grad((want * w).sum(), x64)
    x = img.to(dev()).requires_grad_(True)
    got = log_compress(x)
    np.testing.assert_allclose(got.detach().cpu().numpy(), want.detach().numpy(), rtol=1e-5, atol=1e-7)
    (got * w.float().to(dev())).sum().backward()
    assert_grad_close(x.grad.cpu().numpy(), gw.numpy(), "d log_compress / d img")


def test_adam_step_matches_torch_optim():
    from diffus_b200 import ops
    g = load_golden("training_loop.npz")
    for key, kw in (("adam_params", dict(lr=0.01)), ("adam_params_wd", dict(lr=1e-3, betas=(0.8, 0.99), eps=1e-6, weight_decay=0.1))):
        p = torch.tensor(g["adam_p0"], device=dev())
        state = torch.zeros((2 * p.numel() + 1,), device=dev())
        for i in range(g["adam_grads"].shape[0]):
            ops.adam_step(p, torch.tensor(g["adam_grads"][i], device=dev()), state, **kw)
            np.testing.assert_allclose(p.cpu().numpy(), g[key][i], rtol=2e-6, atol=2e-7)
        assert state[-1].item() == g["adam_grads"].shape[0]


def _training_scene(n=24):
    from diffus_b200 import ImpedanceEstimator
    from diffus_b200.phantoms import mri_phantom, pose_sweep
    torch.manual_seed(5)
    model = ImpedanceEstimator(1)
    with torch.no_grad():
        model.model[4].bias.fill_(1.5)
        model.model[4].weight.mul_(0.3)
    mri = mri_phantom(n, "t2", seed=2) / 1000.0
    sources, dirs = pose_sweep(6, n_rays=5, n=n, seed=9)
    targets = 0.01 * torch.randn((6, 5, 40), generator=torch.Generator().manual_seed(3))
    return model, mri, sources, dirs, targets


@pytest.mark.parametrize("sampler", ["trilinear", "nearest"])
def test_fused_trainer_matches_autograd_and_torch_adam(sampler):
    """FusedTrainer (flat buffers, fused Adam) == train_step (autograd + torch.optim.Adam) == fp64 oracle for the first gradient."""
    import copy
    from diffus_b200.training import FusedTrainer, train_step
    from oracle import port
    model, mri, sources, dirs, targets = _training_scene()
    S, alpha = 40, 1e-3
    ref = copy.deepcopy(model).double()
    Z64 = ref.model(mri.double().reshape(-1, 1)).reshape(mri.shape) * 1e6
    f64 = torch.stack([port.plot_beam_frame(Z64, sources[p].double(), dirs[p].double(), S, alpha, sampler=sampler)[3] for p in range(6)])
    l64 = (f64 - targets.double()).square().mean()
    l64.backward()
    want = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
    m1 = copy.deepcopy(model).to(dev())
    tr = FusedTrainer(m1, mri.to(dev()), lr=1e-2, sampler=sampler, out_scale=1e6)
    mt = copy.deepcopy(model).to(dev())
    tt = FusedTrainer(mt, mri.to(dev()), lr=1e-2, sampler=sampler, out_scale=1e6, gather="texture")
    m2 = copy.deepcopy(model).to(dev())
    opt = torch.optim.Adam(m2.parameters(), lr=1e-2)
    args = (sources.to(dev()), dirs.to(dev()), targets.to(dev()), S, alpha)
    l1 = tr.step(*args)
    assert_grad_close(tr.grads.cpu().numpy(), want.numpy(), f"FusedTrainer weight gradient ({sampler})")
    lt = tt.step(*args)                                        # the same step gathering through the texture unit
    assert_grad_close(tt.grads.cpu().numpy(), want.numpy(), f"FusedTrainer weight gradient ({sampler}, texture gathers)")
    np.testing.assert_allclose(lt.item(), l64.item(), rtol=1e-4)
    np.testing.assert_allclose(l1.item(), l64.item(), rtol=1e-4)
    l2 = train_step(m2, opt, mri.to(dev()), *args, out_scale=1e6, sampler=sampler)
    np.testing.assert_allclose(l1.item(), l2.item(), rtol=1e-5)
    for _ in range(5):
        l1 = tr.step(*args)
        l2 = train_step(m2, opt, mri.to(dev()), *args, out_scale=1e6, sampler=sampler)
    np.testing.assert_allclose(l1.item(), l2.item(), rtol=1e-3)
    for (name, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):       # the module's tensors are views of the flat vector
        np.testing.assert_allclose(a.detach().cpu().numpy(), b.detach().cpu().numpy(), rtol=2e-3, atol=2e-5, err_msg=name)
    # the same six Adam steps by torch in float64 on the oracle's renderer: the loss trajectory, not only the first gradient
    opt64 = torch.optim.Adam(ref.parameters(), lr=1e-2)
    ref.zero_grad()
    for _ in range(6):
        Z = ref.model(mri.double().reshape(-1, 1)).reshape(mri.shape) * 1e6
        f = torch.stack([port.plot_beam_frame(Z, sources[p].double(), dirs[p].double(), S, alpha, sampler=sampler)[3] for p in range(6)])
        l = (f - targets.double()).square().mean()
        opt64.zero_grad()
        l.backward()
        opt64.step()
    np.testing.assert_allclose(l1.item(), l.item(), rtol=5e-2)     # Adam at lr 1e-2 amplifies float32 rounding of near-zero gradients


def test_slice_mode_training_forward_and_trainer():
    """ImpedanceLearner.training_forward's slice mode (GPU notebook cell 16): Z_vol = x.clone(); Z_vol[:, :, k] = mlp(x[:, :, k])."""
    import copy
    from diffus_b200 import UltrasoundRenderer, generate_cone_directions
    from diffus_b200.training import FusedTrainer, training_forward
    from oracle import port
    model, mri, _, _, _ = _training_scene()
    mri = mri + 1.0                                             # the raw MRI is the impedance outside the slice: keep it positive
    k, S, alpha = 11, 40, 1e-3
    src = torch.tensor([12.0, 0.0, float(k)])
    dirs = generate_cone_directions([0.1, 1.0], 0.9, 7)          # the fan lies in the slice p2 = k
    w = None

    def slice_oracle(dt):
        nonlocal w
        r_ = copy.deepcopy(model).to(dt)
        zs = r_.model(mri[:, :, k].to(dt).reshape(-1, 1)).reshape(mri.shape[0], mri.shape[1])
        Z = mri.to(dt).clone()
        Z[:, :, k] = zs
        f_ = port.plot_beam_frame(Z, src.to(dt), dirs.to(dt), S, alpha)[3]
        if w is None:
            w = torch.randn(f_.shape, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
        (f_ * w.to(dt)).sum().backward()
        return r_, f_
    ref, f64 = slice_oracle(torch.float64)
    ref32, _ = slice_oracle(torch.float32)          # the same arithmetic in float32: the drift the 1e-4 may be widened by
    m = copy.deepcopy(model).to(dev())
    x, y, z, frame = training_forward(m, UltrasoundRenderer(S, alpha), mri.to(dev()), src.to(dev()), dirs.to(dev()), slice_idx=k)
    assert_frame_close(frame.detach().cpu().numpy(), f64.detach().numpy(), "slice-mode frame")
    assert (z == k).all()
    (frame * w.float().to(dev())).sum().backward()
    for (name, p), (_, q), (_, q32) in zip(m.named_parameters(), ref.named_parameters(), ref32.named_parameters()):
        assert_grad_close(p.grad.cpu().numpy(), q.grad.numpy(), f"slice mode d/d{name}",
                          noise=float((q32.grad.double() - q.grad).abs().max()))
    # the fused trainer in slice mode: first gradient vs the oracle's MSE gradient
    target = 0.5 * f64.detach().float()
    ref2 = copy.deepcopy(model).double()
    zs2 = ref2.model(mri[:, :, k].double().reshape(-1, 1)).reshape(mri.shape[0], mri.shape[1])
    Z2 = mri.double().clone()
    Z2[:, :, k] = zs2
    l64 = (port.plot_beam_frame(Z2, src.double(), dirs.double(), S, alpha)[3] - target.double()).square().mean()
    l64.backward()
    want = torch.cat([p.grad.reshape(-1) for p in ref2.parameters()])
    tr = FusedTrainer(copy.deepcopy(model).to(dev()), mri.to(dev()), lr=1e-3, sampler="nearest", slice_index=k)
    loss = tr.step(src.to(dev()).reshape(1, 3), dirs.to(dev()), target.to(dev()).unsqueeze(0), S, alpha)
    np.testing.assert_allclose(loss.item(), l64.item(), rtol=1e-4)
    assert_grad_close(tr.grads.cpu().numpy(), want.numpy(), "slice-mode FusedTrainer weight gradient")
