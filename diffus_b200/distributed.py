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


class FlatBuffer:
    """One persistent flat float32 buffer whose views REPLACE the given tensors' storage.

    Whatever writes those tensors (autograd accumulating into ``p.grad``, a kernel handed ``view.data_ptr()``) writes
    straight into the buffer, and the collective runs on the buffer itself -- no ``cat`` before and no ``copy_`` after."""

    def __init__(self, tensors: Sequence[torch.Tensor]):
        tensors = list(tensors)
        n = sum(t.numel() for t in tensors)
        self.flat = torch.zeros((n,), dtype=torch.float32, device=tensors[0].device)
        self.views = []
        off = 0
        for t in tensors:
            v = self.flat[off:off + t.numel()].view(t.shape)
            v.copy_(t)
            t.data = v                       # the caller's tensor now lives inside the flat buffer
            self.views.append(v)
            off += t.numel()

    def allreduce(self, weight: float = 1.0, group=None) -> torch.Tensor:
        _, w = world()
        if weight != 1.0:
            self.flat.mul_(weight)
        if w > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        return self.flat


_module_buffers = {}       # id(module) -> FlatBuffer over its parameters' .grad tensors


def allreduce_grads(tensors: Sequence[torch.Tensor], average: bool = False, group=None, weight: float = 1.0) -> None:
    """``tensors <- sum over ranks of weight * tensors`` (divided by the world size if ``average``) IN PLACE, one collective.
    One-off form for arbitrary tensors; :func:`allreduce_module_grads` keeps a persistent flat buffer instead."""
    _, w = world()
    tensors = [t for t in tensors if t is not None]
    if not tensors:
        return
    scale = weight / (w if average else 1)
    if w == 1:
        if scale != 1.0:
            for t in tensors:
                t.mul_(scale)
        return
    flat = torch.cat([t.reshape(-1).to(torch.float32) for t in tensors])
    if scale != 1.0:
        flat *= scale
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n


def allreduce_module_grads(module: torch.nn.Module, average: bool = False, group=None, weight: Optional[float] = None) -> None:
    """All-reduce ``p.grad`` of every parameter of ``module`` through ONE persistent flat buffer.

    The first call re-homes the ``.grad`` tensors as views of that buffer (missing grads count as zero); autograd then
    accumulates into the views in place, so later calls run the collective on the buffer as it stands.  ``weight`` scales
    this rank's gradients first (its share of the global batch: ragged shards stay exact); ``average`` divides by the
    world size instead (equal shards)."""
    _, w = world()
    if weight is None:
        weight = 1.0 / w if average else 1.0
    params = [p for p in module.parameters() if p.requires_grad]
    buf = _module_buffers.get(id(module))
    stale = buf is None or len(buf.views) != len(params) or any(
        p.grad is None or p.grad.data_ptr() != v.data_ptr() for p, v in zip(params, buf.views))
    if stale:
        for p in params:
            if p.grad is None:
                p.grad = torch.zeros_like(p, dtype=torch.float32)
        buf = FlatBuffer([p.grad for p in params])
        _module_buffers[id(module)] = buf
    buf.allreduce(weight=float(weight), group=group)


def global_share(n_local: float, device) -> float:
    """This rank's fraction of the global batch (elements), for weighting per-rank mean losses and their gradients."""
    _, w = world()
    if w == 1:
        return 1.0
    t = torch.tensor([float(n_local)], dtype=torch.float64, device=device)
    dist.all_reduce(t)
    return float(n_local) / float(t.item())


def gather_frames(local_frames: torch.Tensor, n_poses: Optional[int] = None, group=None) -> torch.Tensor:
    """All-gather pose-sharded frames (P_local, R, S) into (P, R, S) on every rank.  Ragged shards are allowed: with
    ``n_poses`` the sizes follow :func:`pose_shard`; without it the ranks first exchange their shard sizes."""
    rank, w = world()
    if w == 1:
        return local_frames
    if n_poses is not None:
        sizes = shard_sizes(n_poses, w)
    else:
        mine = torch.tensor([local_frames.shape[0]], dtype=torch.int64, device=local_frames.device)
        every = torch.empty((w,), dtype=torch.int64, device=local_frames.device)
        dist.all_gather_into_tensor(every, mine, group=group)
        sizes = [int(v) for v in every.tolist()]
    if len(set(sizes)) == 1:
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
