"""Multi-GPU plumbing for the renderer: one process per GPU, poses sharded, volume replicated.

Every probe pose (frame) is independent, so the pose sweep shards across ranks with NO
data-path collective; each rank holds a full volume replica (64 MiB for 256^3).
Collectives appear only where the path has a real exchange step:

* ``allreduce_grads``   -- gradients of parameters SHARED by all poses (the 1 153 MLP weights,
  or a shared probe pose) summed over ranks in ONE fused flat buffer (4.6 KB: latency-bound);
* ``gather_frames``     -- optional all-gather of rendered frames;
* ``broadcast_volume``  -- once per volume.

The functions take the tensors' device as it is (NCCL for CUDA tensors, gloo for the CPU
tests of the host logic) and are no-ops for a single process.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size) -- (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def pose_shard(n_poses: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> slice:
    """Contiguous block of poses owned by ``rank``; the first ``n % world`` ranks get one extra pose."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(n_poses, world_size)
    lo = rank * base + min(rank, extra)
    return slice(lo, lo + base + (1 if rank < extra else 0))


def shard_sizes(n_poses: int, world_size: int) -> List[int]:
    return [pose_shard(n_poses, r, world_size).stop - pose_shard(n_poses, r, world_size).start
            for r in range(world_size)]


def allreduce_grads(tensors: Sequence[torch.Tensor], average: bool = False, group=None) -> None:
    """Sum (or average) ``tensors`` over ranks IN PLACE with a single collective on a flat buffer."""
    _, w = world()
    tensors = [t for t in tensors if t is not None]
    if w == 1 or not tensors:
        return
    flat = torch.cat([t.reshape(-1).to(torch.float32) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= w
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n


def allreduce_module_grads(module: torch.nn.Module, average: bool = True, group=None) -> None:
    """All-reduce ``p.grad`` of every parameter of ``module`` (missing grads count as zero)."""
    grads = []
    for p in module.parameters():
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        grads.append(p.grad)
    allreduce_grads(grads, average=average, group=group)


def gather_frames(local_frames: torch.Tensor, n_poses: Optional[int] = None, group=None) -> torch.Tensor:
    """All-gather pose-sharded frames (P_local, R, S) into (P, R, S) on every rank (ragged shards allowed)."""
    rank, w = world()
    if w == 1:
        return local_frames
    sizes = shard_sizes(n_poses, w) if n_poses is not None else None
    if sizes is None or len(set(sizes)) == 1:
        out = local_frames.new_empty((local_frames.shape[0] * w,) + tuple(local_frames.shape[1:]))
        dist.all_gather_into_tensor(out, local_frames.contiguous(), group=group)
        return out
    pad = max(sizes)
    buf = local_frames.new_zeros((pad,) + tuple(local_frames.shape[1:]))
    buf[: local_frames.shape[0]] = local_frames
    out = local_frames.new_empty((pad * w,) + tuple(local_frames.shape[1:]))
    dist.all_gather_into_tensor(out, buf, group=group)
    return torch.cat([out[r * pad: r * pad + sizes[r]] for r in range(w)])


def broadcast_volume(volume: torch.Tensor, src: int = 0, group=None) -> torch.Tensor:
    _, w = world()
    if w > 1:
        dist.broadcast(volume, src=src, group=group)
    return volume


def global_mean_loss(local_loss: torch.Tensor, local_count: int, group=None) -> torch.Tensor:
    """Mean over all ranks' elements from per-rank means (shards may be ragged)."""
    _, w = world()
    if w == 1:
        return local_loss
    acc = torch.stack([local_loss.detach().to(torch.float32) * local_count,
                       torch.tensor(float(local_count), device=local_loss.device)])
    dist.all_reduce(acc, group=group)
    return acc[0] / acc[1]
