"""CPU, world_size 2 over gloo: the host-side multi-GPU logic (pose sharding, fused gradient all-reduce,
frame gather).  The rendering itself has no CPU path, so ranks exchange stand-in tensors."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_poses, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from diffus_b200 import distributed as D
        from diffus_b200.impedance import ImpedanceEstimator
        assert D.world() == (rank, world)
        sl = D.pose_shard(n_poses)
        # every pose is rendered exactly once: frame p is filled with the value p
        local = torch.stack([torch.full((3, 5), float(p)) for p in range(sl.start, sl.stop)]) if sl.stop > sl.start \
            else torch.zeros((0, 3, 5))
        frames = D.gather_frames(local, n_poses)
        ok_gather = frames.shape == (n_poses, 3, 5) and all(bool((frames[p] == p).all()) for p in range(n_poses))
        # shared-parameter gradients: one fused all-reduce
        torch.manual_seed(0)
        model = ImpedanceEstimator(1)
        for i, p in enumerate(model.parameters()):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        D.allreduce_module_grads(model, average=True)
        want = [sum(r + 1 for r in range(world)) / world * (i + 1) for i in range(6)]
        ok_grads = all(bool(torch.allclose(p.grad, torch.full_like(p, w))) for p, w in zip(model.parameters(), want))
        # second step: autograd-style in-place accumulation lands in the SAME persistent flat buffer (no cat / copy_ back)
        ptrs = [p.grad.data_ptr() for p in model.parameters()]
        for i, p in enumerate(model.parameters()):
            p.grad.zero_()
            p.grad.add_(float(rank + 2) * (i + 1))
        share = D.global_share(sl.stop - sl.start, torch.device("cpu"))
        D.allreduce_module_grads(model, weight=share)
        sizes = D.shard_sizes(n_poses, world)
        want2 = [sum((r + 2) * sizes[r] for r in range(world)) / n_poses * (i + 1) for i in range(6)]
        ok_grads = ok_grads and ptrs == [p.grad.data_ptr() for p in model.parameters()] and \
            all(bool(torch.allclose(p.grad, torch.full_like(p, w))) for p, w in zip(model.parameters(), want2))
        # ragged shards without n_poses: the ranks exchange their sizes first
        frames2 = D.gather_frames(local)
        ok_gather = ok_gather and frames2.shape == (n_poses, 3, 5) and all(bool((frames2[p] == p).all()) for p in range(n_poses))
        shared = [torch.tensor([1.0, 2.0, 3.0]) * (rank + 1), None, torch.ones(2, 2) * rank]
        D.allreduce_grads(shared)
        ok_sum = bool(torch.allclose(shared[0], torch.tensor([1.0, 2.0, 3.0]) * sum(r + 1 for r in range(world)))) and \
            bool(torch.allclose(shared[2], torch.ones(2, 2) * sum(range(world))))
        loss = D.global_mean_loss(torch.tensor(float(rank + 1)), sl.stop - sl.start)
        sizes = D.shard_sizes(n_poses, world)
        want_loss = sum((r + 1) * sizes[r] for r in range(world)) / n_poses
        vol = torch.full((4, 4, 4), float(rank))
        D.broadcast_volume(vol, src=0)
        q.put((rank, ok_gather, ok_grads, ok_sum, abs(loss.item() - want_loss) < 1e-6, bool((vol == 0).all())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_poses", [8, 7])
def test_world_size_2_gloo(n_poses):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_poses, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in results:
        assert all(r[1:]), f"rank {r[0]}: gather/grads/sum/loss/broadcast = {r[1:]}"


def test_pose_shard_partition_properties():
    from diffus_b200.distributed import pose_shard, shard_sizes
    for n in (0, 1, 5, 8, 1024, 1027):
        for w in (1, 2, 3, 8):
            slices = [pose_shard(n, r, w) for r in range(w)]
            assert slices[0].start == 0 and slices[-1].stop == n
            assert all(a.stop == b.start for a, b in zip(slices, slices[1:]))
            sizes = shard_sizes(n, w)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        pose_shard(4, 2, 2)
    assert pose_shard(10) == slice(0, 10)                  # not initialised: single process
