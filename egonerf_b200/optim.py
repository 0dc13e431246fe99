"""Table-space optimiser.

The reference runs torch.optim.Adam over the 24 factor tensors (train.py:172-186) and then refreshes the pooled coarse
grid (train.py:356-357).  With the render tables as the kernels' view of those tensors, a step would be: transpose the
table-layout gradient to NCHW (unpack), Adam, re-interleave into fp32 tables, convert to bf16 / fp16, pool — five passes over
~100 MB.  `TableAdam` does the same arithmetic in ONE pass in table space (`egn_adam_tables`): it consumes the gradient
exactly as `egn_render_backward` produced it, keeps the moments in table layout and writes NCHW parameters (so
`state_dict()` / checkpoints stay the reference's), fp32 tables, bf16 tables, half tables and coarse tables.  Everything that
is not a factor tensor (basis matrices, MLP, envmap) is handed to torch's fused Adam.

    opt = TableAdam(model, lr_spatial=0.02, lr_network=0.001)      # instead of torch.optim.Adam(model.get_optparam_groups())
    loss.backward()
    opt.regularize(tv_density=w_d, tv_app=w_a, l1_density=w_l1)     # train.py:288-305, fused in table space (optional)
    opt.step(); opt.zero_grad()
    for g in opt.param_groups: g['lr'] *= lr_factor                 # train.py:328-329 works unchanged

Gradients from any OTHER loss on the factor Parameters (plain-torch `model.TV_loss_density(reg)`, `vector_comp_diffs()`, ...)
arrive in `p.grad` through autograd; `step()` folds them into the table-layout gradient (`egn_pack_table_grads`) before the
update and `zero_grad()` clears them, so nothing is dropped or accumulated across steps.
"""
from __future__ import annotations

import torch

from . import _lib


class TableAdam:
    def __init__(self, model, lr_spatial=0.02, lr_network=0.001, lr_envmap=0.1, betas=(0.9, 0.99), eps=1e-8):
        self.model = model
        self.betas, self.eps = betas, eps
        self.step_count = 0
        self.d_tables = None
        tables = model._render_tables()
        self.exp_avg = torch.zeros_like(tables)
        self.exp_avg_sq = torch.zeros_like(tables)
        factor_ids = {id(p) for p in model._factor_params()}
        rest = [{'params': [p for p in g['params'] if id(p) not in factor_ids], 'lr': g['lr']}
                for g in model.get_optparam_groups(lr_spatial, lr_network, lr_envmap, merged=True)]
        rest = [g for g in rest if g['params']]
        self.other = torch.optim.Adam(rest, betas=betas, eps=eps, fused=True) if rest else None
        self.factor_group = {'lr': lr_spatial, 'params': [], 'name': 'factor tensors (table space)'}
        self.param_groups = [self.factor_group] + (self.other.param_groups if self.other else [])
        self.reg_losses = None
        # None: all-reduce the table gradient in fp32 (exact: the sum of the shards' gradients); torch.bfloat16: half the bytes
        # over NVLink at 2^-9 relative rounding per element (a measured option of bench.py --grad-dtype, not the default)
        self.grad_allreduce_dtype = None
        # ray-sharded training over NVLink peer memory (enable_peer_exchange): the gradient buffer then IS peer memory
        self.peer = None
        self._n_tables = tables.numel()
        model._table_opt = self

    def enable_peer_exchange(self, group=None, tail_floats=None, blocks=148):
        """Collective over `group`: from now on the table-layout gradient is accumulated in a buffer every rank can reach over
        NVLink (cudaIpc peer memory, `sharding.PeerExchange`) and `allreduce()` is ONE kernel over peer memory
        (`egn_peer_allreduce`) instead of `ncclAllReduce`; the buffer's tail carries the flat bucket of the basis / MLP
        gradients (`EgoNeRF.allreduce_gradients`), so the step needs no other gradient collective.  Returns False (and keeps
        the NCCL path on every rank) when peer memory cannot be mapped."""
        import torch.distributed as dist
        from .sharding import PeerExchange
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return False
        if tail_floats is None:                      # everything but the factor tensors and the envmap emission
            fac = {id(p) for p in self.model._factor_params()}
            em = id(self.model.envmap.emission) if self.model.envmap is not None else None
            tail_floats = sum(p.numel() for p in self.model._param_list() if id(p) not in fac and id(p) != em)
        self._n_head = (self._n_tables + 3) // 4 * 4
        try:
            self.peer = PeerExchange(self._n_head + int(tail_floats), self._device(), group, blocks)
        except Exception as e:                       # raised on every rank alike (PeerExchange agrees on failure collectively)
            import warnings
            warnings.warn(f"egonerf_b200: {e}; gradient exchange stays on NCCL")
            self.peer = None
            return False
        self.peer_head = self.peer.tensor[:self._n_tables]
        self.peer_tail = self.peer.tensor[self._n_head:]
        self.d_tables = None
        return True

    def disable_peer_exchange(self):
        """Collective: back to the NCCL path; unmaps and frees the peer buffers."""
        if self.peer is not None:
            self.d_tables = None
            bucket = getattr(self.model, "_bucket", None)
            if bucket is not None and bucket.external:           # .grad views into the buffer that is about to be freed
                for p in bucket.params:
                    p.grad = None
            self.peer_head = self.peer_tail = None
            self.peer.close()
            self.peer = None
            self.model._bucket = None

    def grad_target(self, like):
        """Zeroed table-layout buffer for one backward pass: the peer buffer itself for the first backward of a step."""
        if self.peer is not None and self.d_tables is None:
            return self.peer_head.view_as(like).zero_()
        return torch.zeros_like(like)

    def _device(self):
        return self.exp_avg.device

    def _grad_buffer(self):
        if self.d_tables is None:
            self.d_tables = self.grad_target(self.exp_avg)
        return self.d_tables

    # called by the render autograd node instead of egn_unpack_table_grads
    def accumulate(self, d_tables):
        if self.d_tables is None or self.d_tables is d_tables:
            self.d_tables = d_tables
        else:
            self.d_tables.add_(d_tables)

    def zero_grad(self, set_to_none=True):
        self.d_tables = None
        for p in self.model._factor_params():        # gradients of plain-torch losses on the factor Parameters
            p.grad = None
        if self.other:
            self.other.zero_grad(set_to_none=set_to_none)
        bucket = getattr(self.model, "_bucket", None)
        if bucket is not None:                       # ray-sharded training: gradients accumulate in place in the exchange bucket
            bucket.zero()

    @torch.no_grad()
    def regularize(self, tv_density=0.0, tv_app=0.0, l1_density=0.0):
        """Adds the gradients of  tv_density * model.TV_loss_density(TVLoss()) + tv_app * model.TV_loss_app(TVLoss())
        + l1_density * model.density_L1()  (train.py:288-305) to the table-layout gradient in one pass over the render tables
        (`egn_regularize_tables`) and returns the three UNWEIGHTED loss values as a device tensor (for logging)."""
        m = self.model
        with torch.cuda.device(self._device()):
            lib = _lib.load()
            losses = torch.zeros(3, device=self._device(), dtype=torch.float32)
            _lib.check(lib.egn_regularize_tables(m._config(None), m._render_tables().data_ptr(), self._grad_buffer().data_ptr(),
                                                 float(tv_density), float(tv_app), float(l1_density), losses.data_ptr(),
                                                 torch.cuda.current_stream().cuda_stream))
        self.reg_losses = losses
        return losses

    def _fold_param_grads(self):
        """p.grad of the 24 factor tensors (written by autograd for losses outside egn_render_backward) -> d_tables."""
        m = self.model
        fp = m._factor_params()
        grads = [None if p.grad is None else p.grad.detach().contiguous().float() for p in fp]
        if all(g is None for g in grads):
            return
        lib = _lib.load()
        G = m._grads_struct(grads + [None] * (len(m._param_list()) - 24))
        _lib.check(lib.egn_pack_table_grads(m._config(None), G, self._grad_buffer().data_ptr(),
                                            torch.cuda.current_stream().cuda_stream))
        for p in fp:
            p.grad = None

    def allreduce(self, group=None, average=False):
        """Ray-sharded data parallelism: the table-layout factor gradient is one contiguous buffer already.  A rank that ran
        no backward this step still takes part (with zeros): skipping the collective would hang the other ranks."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        self._fold_param_grads()
        buf = self._grad_buffer()
        if self.peer is not None:
            if buf.data_ptr() != self.peer_head.data_ptr():      # gradient was produced outside the peer buffer
                self.peer_head.view_as(buf).copy_(buf)
                buf = self.d_tables = self.peer_head.view_as(buf)
            # head (factor gradient) and tail (basis / MLP bucket) in one kernel; the mean over ranks is folded in
            self.peer.allreduce(1.0 / dist.get_world_size(group) if average else 1.0)
            return
        if self.grad_allreduce_dtype is not None:
            low = buf.to(self.grad_allreduce_dtype)
            dist.all_reduce(low, op=dist.ReduceOp.SUM, group=group)
            buf.copy_(low)
        else:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        if average:
            buf.div_(dist.get_world_size(group))

    @torch.no_grad()
    def step(self):
        m = self.model
        with torch.cuda.device(self._device()):
            self._fold_param_grads()
            if self.d_tables is not None:
                lib = _lib.load()
                self.step_count += 1
                cfg = m._config(None)
                plist = m._param_list()
                P = m._grads_struct([p.detach() for p in plist])        # destinations: the NCHW parameter tensors themselves
                tables = m._render_tables()                              # current tables (== current parameters)
                t16 = m._tables_bf16.data_ptr() if m._tables_bf16 is not None else None
                th = m._tables_h.data_ptr() if getattr(m, "_tables_h", None) is not None else None
                _lib.check(lib.egn_adam_tables(cfg, P, self.d_tables.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                               tables.data_ptr(), t16, th, float(self.factor_group['lr']), float(self.betas[0]),
                                               float(self.betas[1]), float(self.eps), self.step_count,
                                               torch.cuda.current_stream().cuda_stream))
                # parameters were written through raw pointers: their torch versions did not move, so the (data_ptr, version)
                # key of `_render_tables` still matches and the tables the kernel just refreshed stay in use
            if self.other:
                self.other.step()
