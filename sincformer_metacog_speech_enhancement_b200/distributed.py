"""Data-parallel plumbing for the SincConv path: the batch is sharded across ranks (one process per
GPU), forward and grad-input need no communication, and the only exchange is one all-reduce of the
2*C-element parameter gradient (512 bytes at C=64) -- SURVEY.md section 8(e).

Under torch's DistributedDataParallel the two parameters ride DDP's bucket and this helper must NOT be
called as well.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_rows(global_batch: int, rank: int, world_size: int) -> tuple[int, int]:
    """Rows [begin, end) of the global batch owned by `rank` (contiguous, sizes differ by at most 1)."""
    base, rem = divmod(global_batch, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allreduce_filter_grads(module, group=None, average: bool = False) -> None:
    """SUM (or mean) all-reduce of low_hz_.grad / band_hz_.grad as ONE flat 2*C fp32 buffer."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    params = [module.low_hz_, module.band_hz_]
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
    flat = torch.cat([g.reshape(-1).float() for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for p, g in zip(params, grads):
        n = g.numel()
        new = flat[off:off + n].view_as(g).to(g.dtype)
        if p.grad is None:
            p.grad = new
        else:
            p.grad.copy_(new)
        off += n
