"""GPU, needs >= 2 devices (skipped otherwise): pose sharding + NCCL against the single-GPU result.

Two ranks render disjoint pose shards of one sweep and (a) all-gather the frames, (b) all-reduce the
gradients of parameters shared by all poses (the MLP weights) -- the only two collectives of the path.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _scene():
    from diffus_b200 import ImpedanceEstimator
    from diffus_b200.phantoms import mri_phantom, pose_sweep
    torch.manual_seed(11)
    model = ImpedanceEstimator(1)
    with torch.no_grad():
        model.model[4].bias.fill_(1.5)
        model.model[4].weight.mul_(0.3)
    mri = mri_phantom(32, "t2", seed=1) / 1000.0
    sources, dirs = pose_sweep(6, n_rays=8, n=32, seed=4)
    targets = 0.01 * torch.randn((6, 8, 48), generator=torch.Generator().manual_seed(2))
    return model, mri, sources, dirs, targets


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from diffus_b200 import distributed as D, render_frames
        from diffus_b200.training import mlp_render_mse_loss
        model, mri, sources, dirs, targets = _scene()
        model = model.to(dev)
        sl = D.pose_shard(sources.shape[0])
        vol = model.impedance_volume(mri.to(dev), None, 1e6, 400.0).detach()
        D.broadcast_volume(vol, src=0)
        frames = D.gather_frames(render_frames(vol, sources[sl].to(dev), dirs[sl].to(dev), 48, 1e-3, sampler="trilinear"),
                                 sources.shape[0])
        loss = mlp_render_mse_loss(model, mri.to(dev), sources[sl].to(dev), dirs[sl].to(dev), targets[sl].to(dev), 48, 1e-3,
                                   out_scale=1e6)
        loss.backward()
        D.allreduce_module_grads(model, average=True)
        gl = D.global_mean_loss(loss.detach(), targets[sl].numel())
        q.put((rank, frames.cpu(), [p.grad.cpu() for p in model.parameters()], gl.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_match_single_gpu():
    from diffus_b200 import render_frames
    from diffus_b200.training import mlp_render_mse_loss
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=300) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    dev = torch.device("cuda:0")
    model, mri, sources, dirs, targets = _scene()
    model = model.to(dev)
    vol = model.impedance_volume(mri.to(dev), None, 1e6, 400.0).detach()
    frames = render_frames(vol, sources.to(dev), dirs.to(dev), 48, 1e-3, sampler="trilinear").cpu()
    loss = mlp_render_mse_loss(model, mri.to(dev), sources.to(dev), dirs.to(dev), targets.to(dev), 48, 1e-3, out_scale=1e6)
    loss.backward()
    for rank, fr, grads, gl in results:
        assert torch.equal(fr, frames), "gathered frames differ from the single-GPU sweep"
        assert abs(gl - loss.item()) <= 1e-5 * abs(loss.item())
        for g, p in zip(grads, model.parameters()):
            scale = p.grad.abs().max().item() + 1e-30
            assert (g - p.grad.cpu()).abs().max().item() <= 2e-4 * scale
