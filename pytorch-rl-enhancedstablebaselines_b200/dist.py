"""Multi-GPU plumbing: one process per GPU, reactors sharded by contiguous global id ranges.

Reactor instances never interact (the reference steps them one after another,
``core/common/vec_env/dummy_vec_env.py:58-72``), so the rollout needs NO collective: rank r owns global
ids ``[offset, offset + count)``, its own replay shard and — because every Philox counter is keyed by
the *global* reactor id — produces exactly the rows a single-GPU run would produce for those ids.
The only exchange on the path is the data-parallel gradient all-reduce of the shared policy update
(TD3: 369,704 fp32 = 1.48 MB): one flat bucket, one ``all_reduce`` (NCCL over NVLink on GPUs, gloo in
the CPU tests).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple


def shard_range(n_total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """(offset, count) of rank's contiguous shard; the first ``n_total % world_size`` ranks get one extra."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, extra = divmod(n_total, world_size)
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def make_sharded_env(n_total: int, rank: int, world_size: int, device=None, **env_kwargs):
    """This rank's ``GpuCSTRVecEnv`` shard of a job with ``n_total`` reactors."""
    from .env import GpuCSTRVecEnv

    offset, count = shard_range(n_total, rank, world_size)
    if device is None:
        device = f"cuda:{rank}"
    return GpuCSTRVecEnv(count, device=device, env_offset=offset, **env_kwargs)


def _flat_views(tensors: List, flat) -> List:
    out, pos = [], 0
    for t in tensors:
        n = t.numel()
        out.append(flat[pos:pos + n].view_as(t))
        pos += n
    return out


def allreduce_gradients(params: Iterable, group=None, average: bool = True, bucket=None):
    """Sum (or average) the ``.grad`` of ``params`` over the process group with ONE collective on a flat
    fp32 bucket.  Returns the bucket so callers can reuse it (``bucket=`` on the next call)."""
    import torch
    import torch.distributed as dist

    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return bucket
    total = sum(g.numel() for g in grads)
    if bucket is None or bucket.numel() != total or bucket.device != grads[0].device:
        bucket = torch.empty(total, dtype=torch.float32, device=grads[0].device)
    views = _flat_views(grads, bucket)
    torch._foreach_copy_(views, grads)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
        if average:
            bucket.div_(dist.get_world_size(group))
    torch._foreach_copy_(grads, views)
    return bucket


def allreduce_flat(flat, group=None, average: bool = True):
    """In-place all-reduce (mean by default) of an already flat gradient range — the ``allreduce=`` hook of
    ``FusedTD3Update.update``: the flat ``grads`` block IS the NCCL bucket, no packing copies."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if average and flat.is_cuda:  # NCCL averages inside the collective: no extra scaling launch (and none to capture in a graph)
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            if average:
                flat.div_(dist.get_world_size(group))
    return flat


def broadcast_parameters(params: Iterable, src: int = 0, group=None) -> None:
    """Make every rank start from rank ``src``'s parameters (one flat broadcast)."""
    import torch
    import torch.distributed as dist

    ps = [p.data for p in params]
    if not ps or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([p.reshape(-1).float() for p in ps])
    dist.broadcast(flat, src=src, group=group)
    torch._foreach_copy_(ps, _flat_views(ps, flat))


def global_sum(value: float, device=None, group=None) -> float:
    """Sum of a Python scalar over ranks (episode statistics, transitions counters)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, group=group)
    return float(t.item())
