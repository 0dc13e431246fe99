"""Table-space optimiser.

The reference runs torch.optim.Adam over the 24 factor tensors (train.py:172-186) and then refreshes the pooled coarse
grid (train.py:356-357).  With the render tables as the kernels' view of those tensors, a step would be: transpose the
table-layout gradient to NCHW (unpack), Adam, re-interleave into fp32 tables, convert to bf16, pool — five passes over
~100 MB.  `TableAdam` does the same arithmetic in ONE pass in table space (`egn_adam_tables`): it consumes the gradient
exactly as `egn_render_backward` produced it, keeps the moments in table layout and writes NCHW parameters (so
`state_dict()` / checkpoints stay the reference's), fp32 tables, bf16 tables and coarse tables.  Everything that is not a
factor tensor (basis matrices, MLP, envmap) is handed to torch's fused Adam.

    opt = TableAdam(model, lr_spatial=0.02, lr_network=0.001)      # instead of torch.optim.Adam(model.get_optparam_groups())
    loss.backward(); opt.step(); opt.zero_grad()
    for g in opt.param_groups: g['lr'] *= lr_factor                 # train.py:328-329 works unchanged
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
        self.tables_fresh = False
        tables = model._render_tables()
        self.exp_avg = torch.zeros_like(tables)
        self.exp_avg_sq = torch.zeros_like(tables)
        rest = [g for g in model.get_optparam_groups(lr_spatial, lr_network, lr_envmap, merged=True) if g['lr'] != lr_spatial
                or not any(p is q for p in g['params'] for q in model._factor_params())]
        factor_ids = {id(p) for p in model._factor_params()}
        rest = [{'params': [p for p in g['params'] if id(p) not in factor_ids], 'lr': g['lr']} for g in rest]
        rest = [g for g in rest if g['params']]
        self.other = torch.optim.Adam(rest, betas=betas, eps=eps, fused=True) if rest else None
        self.factor_group = {'lr': lr_spatial, 'params': [], 'name': 'factor tensors (table space)'}
        self.param_groups = [self.factor_group] + (self.other.param_groups if self.other else [])
        model._table_opt = self

    # called by the render autograd node instead of egn_unpack_table_grads
    def accumulate(self, d_tables):
        self.d_tables = d_tables if self.d_tables is None else self.d_tables.add_(d_tables)

    def zero_grad(self, set_to_none=True):
        self.d_tables = None
        if self.other:
            self.other.zero_grad(set_to_none=set_to_none)

    def allreduce(self, group=None, average=False):
        """Ray-sharded data parallelism: the table-layout factor gradient is one contiguous buffer already."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1 or self.d_tables is None:
            return
        dist.all_reduce(self.d_tables, op=dist.ReduceOp.SUM, group=group)
        if average:
            self.d_tables.div_(dist.get_world_size(group))

    @torch.no_grad()
    def step(self):
        m = self.model
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
            self.tables_fresh = True        # parameters were written through raw pointers: torch versions did not move
        if self.other:
            self.other.step()
