"""Device-resident batch assembly.

The reference keeps `allrays (N,6)` / `allrgbs (N,3)` on the host, draws a permutation with numpy
(`sampler.SimpleSampler`, sampler.py:4-16), fancy-indexes on the CPU and copies every batch to the device
(train.py:247-248, renderer.py:26).  At millions of rays per second that host gather is the bottleneck; here the ray
bank lives in HBM, the epoch permutation is drawn on the device and a batch is one device-side index_select.
ERP frames are generated on the device from the 3 x 4 pose (`erp_rays`) instead of being shipped as 24 B/ray."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class RayBank:
    """Same iteration semantics as SimpleSampler: a fresh permutation of all rays whenever fewer than `batch` unseen
    rays remain; consecutive batches of one permutation are disjoint."""

    def __init__(self, all_rays: torch.Tensor, all_rgbs: torch.Tensor, batch: int, device="cuda", seed: int | None = None,
                 rank: int = 0, world: int = 1):
        from .sharding import shard_range
        a, b = shard_range(all_rays.shape[0], rank, world)            # ray-sharded data parallelism: disjoint blocks
        self.rays = all_rays[a:b].to(device, non_blocking=True).contiguous().float()
        self.rgbs = all_rgbs[a:b].to(device, non_blocking=True).contiguous().float()
        self.total, self.batch = self.rays.shape[0], int(batch)
        self.curr = self.total
        self.ids = None
        self.gen = torch.Generator(device=self.rays.device)
        if seed is not None:
            self.gen.manual_seed(seed + rank)

    def nextids(self) -> torch.Tensor:
        self.curr += self.batch
        if self.curr + self.batch > self.total:
            self.ids = torch.randperm(self.total, device=self.rays.device, generator=self.gen)
            self.curr = 0
        return self.ids[self.curr:self.curr + self.batch]

    def next_batch(self):
        idx = self.nextids()
        return self.rays.index_select(0, idx), self.rgbs.index_select(0, idx)


def erp_rays(H: int, W: int, c2w, rows=None, device="cuda") -> torch.Tensor:
    """(rows * W, 6) rays [origin, unit direction] of an equirectangular frame, generated on the device."""
    lib = _lib.load()
    r0, r1 = rows if rows is not None else (0, H)
    pose = torch.as_tensor(c2w, dtype=torch.float32).reshape(-1)[:12].contiguous().cpu()
    out = torch.empty((r1 - r0) * W, 6, device=device, dtype=torch.float32)
    if not out.is_cuda:
        raise RuntimeError("egonerf_b200: erp_rays generates rays on a CUDA device — there is no CPU fallback")
    buf = (C.c_float * 12)(*pose.tolist())
    with torch.cuda.device(out.device):
        _lib.check(lib.egn_erp_rays(H, W, r0, r1 - r0, buf, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return out
