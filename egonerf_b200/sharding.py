"""Ray-sharded data parallelism: rays partition across ranks, the grid is replicated, one all-reduce
(sum) over the gradients per training step.  The reference is single-process (train.py:20); this is the host-side
plumbing a multi-GPU launcher wraps around the drop-in modules.  Backend: NCCL over NVLink on GPUs, gloo in CPU tests."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_rays: int, rank: int, world: int):
    """Contiguous block [start, stop) of rank `rank`; blocks differ by at most one ray and cover [0, n) exactly."""
    base, rem = divmod(int(n_rays), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def row_tiles(height: int, rank: int, world: int):
    """Row tile of an equirectangular frame for tiled rendering (BASELINE.json configs[4])."""
    return shard_range(height, rank, world)


class GradientBucket:
    """One flat buffer holding every gradient, so that a step needs ONE all-reduce launch (latency-bound over NVSwitch:
    bucket for launch count, not link count).  `p.grad` of every parameter becomes a view into the buffer."""

    def __init__(self, params):
        self.params = [p for p in params]
        total = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.flat = torch.zeros(total, dtype=p0.dtype, device=p0.device)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def attach(self):
        for p, v in zip(self.params, self.views):
            p.grad = v

    def zero(self):
        self.flat.zero_()
        self.attach()

    def gather_from_params(self):
        """Copies gradients that autograd allocated on its own back into the bucket (and re-attaches the views)."""
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
            p.grad = v

    def allreduce(self, group=None, average=False):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            self.flat.div_(dist.get_world_size(group))
