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

    def __init__(self, params, storage=None):
        self.params = [p for p in params]
        total = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        # `storage`: a flat fp32 tensor someone else reduces (the tail of the peer-memory buffer, PeerExchange)
        self.external = storage is not None
        self.flat = storage[:total].zero_() if self.external else torch.zeros(total, dtype=p0.dtype, device=p0.device)
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


def slice_bounds(n_floats: int, rank: int, world: int):
    """[start, stop) in floats of the slice rank `rank` sums in `egn_peer_allreduce` (egn_peer.cu: float4 units, ceil split)."""
    n4 = int(n_floats) // 4
    per = -(-n4 // int(world))
    lo = min(per * rank, n4)
    return 4 * lo, 4 * min(lo + per, n4)


class _DevicePointer:
    """`__cuda_array_interface__` view of device memory this library allocated, so torch can wrap it without owning it."""

    def __init__(self, ptr, numel):
        self.__cuda_array_interface__ = {"shape": (int(numel),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class PeerExchange:
    """Gradient buffer in NVLink peer memory + the one-kernel all-reduce over it (`egn_peer_allreduce`, csrc/egn_peer.cu).

    Every rank allocates `numel` floats (+ a flag block) with `egn_peer_alloc`, the ranks exchange cudaIpc handles through the
    process group and map each other's allocations.  `tensor` is the local buffer as a torch tensor: the backward kernels
    scatter into it, `allreduce()` replaces it by the sum over ranks (bit-identical on every rank) and `egn_adam_tables`
    reads it -- no copy in between.  Construction is collective; if any rank cannot map its peers, all ranks raise and the
    caller keeps the NCCL path."""

    def __init__(self, numel, device, group=None, blocks=148):
        import ctypes as C
        from . import _lib
        self.lib = _lib.load()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device(device)
        self.numel = (int(numel) + 3) // 4 * 4
        self.blocks, self.epoch = int(blocks), 0
        self._opened, self._own = [], []
        ok, err = True, ""
        with torch.cuda.device(self.device):
            try:
                buf, flg = C.c_void_p(), C.c_void_p()
                _lib.check(self.lib.egn_peer_alloc(self.numel * 4, C.byref(buf)))
                self._own.append(buf.value)
                _lib.check(self.lib.egn_peer_alloc(self.lib.egn_peer_flag_bytes(), C.byref(flg)))
                self._own.append(flg.value)
                hb, hf = C.create_string_buffer(64), C.create_string_buffer(64)
                _lib.check(self.lib.egn_peer_export(buf, hb))
                _lib.check(self.lib.egn_peer_export(flg, hf))
                mine = (self.rank, bytes(hb.raw), bytes(hf.raw))
            except RuntimeError as e:                    # still take part in the collectives below
                ok, err, mine = False, str(e), (self.rank, b"", b"")
            handles = [None] * self.world
            dist.all_gather_object(handles, mine, group=group)
            bufs, flags = [None] * self.world, [None] * self.world
            if ok:
                bufs[self.rank], flags[self.rank] = buf.value, flg.value
                try:
                    for r, b, f in handles:
                        if r == self.rank:
                            continue
                        if not b:
                            raise RuntimeError(f"rank {r} could not export its buffer")
                        pb, pf = C.c_void_p(), C.c_void_p()
                        _lib.check(self.lib.egn_peer_open(b, C.byref(pb)))
                        self._opened.append(pb.value)
                        _lib.check(self.lib.egn_peer_open(f, C.byref(pf)))
                        self._opened.append(pf.value)
                        bufs[r], flags[r] = pb.value, pf.value
                except RuntimeError as e:
                    ok, err = False, str(e)
            agree = torch.tensor([1 if ok else 0], device=self.device, dtype=torch.int32)
            dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=group)
            if int(agree.item()) == 0:
                self.close()
                raise RuntimeError("peer-memory exchange unavailable: " + (err or "a peer rank failed to map the buffers"))
            self._bufs = (C.c_void_p * self.world)(*bufs)
            self._flags = (C.c_void_p * self.world)(*flags)
            self._holder = _DevicePointer(buf.value, self.numel)
            self.tensor = torch.as_tensor(self._holder, device=self.device)
            # nobody may start the first exchange before every rank has mapped everything
            dist.barrier(group=group)

    def allreduce(self, scale=1.0, stream=None):
        """tensor <- scale * sum over ranks of tensor, on `stream` (default: torch's current stream)."""
        from . import _lib
        self.epoch += 1
        with torch.cuda.device(self.device):
            st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
            _lib.check(self.lib.egn_peer_allreduce(self._bufs, self._flags, self.rank, self.world, self.numel, float(scale),
                                                   self.epoch, self.blocks, st))

    def close(self):
        """Collective: every rank unmaps its peers' allocations before anybody frees its own."""
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            for p in self._opened:
                self.lib.egn_peer_close(p)
            self._opened = []
            if dist.is_initialized():
                dist.barrier(group=self.group)
            for p in self._own:
                self.lib.egn_peer_free(p)
            self._own = []
        self.tensor = None
