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


def gather_env_gradient(model, group=None, average=False, equal_counts=True):
    """Envmap gradient of a ray-sharded step without the dense all-reduce: the backward pass left, per ray, the gradient
    w.r.t. its env radiance (`model._env_rays`, filled when `model.sparse_env_grad` is set); ranks all-gather directions +
    gradients (24 B / ray; 0.4 MB per rank at 16 384 rays against 88 MB for the dense (3, 3840, 1920) tensor) and every rank
    scatters ALL rays into its own dense gradient with `egn_envmap_backward` -- the same sum, computed locally.
    `equal_counts` (the ray-sharded launcher's contract: every rank steps the same batch size) skips the exchange of the
    per-rank ray counts, which would cost two host synchronisations per step."""
    from . import _lib
    em = model.envmap.emission
    if not model._env_rays:
        packed = torch.zeros(0, 6, device=em.device)
    else:
        packed = torch.cat([torch.cat([d, g], 1) for d, g in model._env_rays], 0).contiguous()
    model._env_rays = []
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    if world > 1:
        if equal_counts:
            padded = packed
        else:
            n_local = torch.tensor([packed.shape[0]], device=em.device, dtype=torch.int64)
            counts = [torch.zeros_like(n_local) for _ in range(world)]
            dist.all_gather(counts, n_local, group=group)
            n_max = max(int(c.item()) for c in counts)
            padded = torch.zeros(n_max, 6, device=em.device)
            padded[:packed.shape[0]] = packed                  # padding rows carry zero gradient: they scatter nothing
        gathered = torch.empty(world * padded.shape[0], 6, device=em.device)
        dist.all_gather_into_tensor(gathered, padded, group=group)
        packed = gathered
    d_em = torch.zeros_like(em)
    if packed.shape[0] > 0:
        lib = _lib.load()
        cfg = _lib.EgnConfig()
        cfg.env_h = em.shape[2]
        dirs, g = packed[:, :3].contiguous(), packed[:, 3:].contiguous()
        with torch.cuda.device(em.device):
            _lib.check(lib.egn_envmap_backward(cfg, em.data_ptr(), dirs.data_ptr(), dirs.shape[0], g.data_ptr(), d_em.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream))
    if average:
        d_em.div_(world)
    em.grad = d_em
