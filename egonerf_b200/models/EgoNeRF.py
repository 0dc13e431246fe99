"""Drop-in `EgoNeRF` module: the reference's operator surface (models/EgoNeRF.py:27-602, models/tensorBase.py:132-268)
on top of libegn_b200.  Parameter names, shapes and state_dict keys are the reference's, so checkpoints and
`train.py`'s optimiser groups work unchanged; `forward` hands raw device pointers to the C ABI.

No CPU path: every compute entry point raises if the tensors are not on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
from math import pi

import numpy as np
import torch
import torch.nn.functional as F

from .. import _lib
from .coordinates import YinYangSphericalCoords, sample_schedule, plain_sample_schedule
from .envmap import EnvironmentMap
from .decoders import MLPRender_Fea, MLPRender, SHRender, RGBRender

MAT_MODE = [[0, 1], [0, 2], [1, 2]]     # EgoNeRF.py:30-33
VEC_MODE = [2, 1, 0]


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _on(t):
    """Every ctypes launch runs under the device of its tensors (the reference supports `device='cuda:1'` while the current
    device is 0): the kernels launch on the CURRENT device's stream, so make the tensors' device current."""
    return torch.cuda.device(t.device)


def _need_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"egonerf_b200: {what} must be a CUDA tensor — there is no CPU fallback")


class YinYangAlphaGridMask(torch.nn.Module):
    """EgoNeRF.py:11-24: one binary occupancy volume per hemisphere, (1,1,N_phi',N_theta',N_r'), sampled trilinearly at
    the active hemisphere's normalised coordinates.  Only `compute_alpha` reads it (the reference's `EgoNeRF.forward` never
    does, SURVEY.md 8a "quirks"); a few torch ops, not on the path."""

    def __init__(self, device, alpha_volume_yin, alpha_volume_yang):
        super().__init__()
        self.device = device
        self.alpha_volume_yin = alpha_volume_yin.view(1, 1, *alpha_volume_yin.shape[-3:])
        self.alpha_volume_yang = alpha_volume_yang.view(1, 1, *alpha_volume_yang.shape[-3:])

    def sample_alpha(self, norm_samples):
        """(M,7) normalised coordinates -> (M,) occupancy of the active hemisphere (column 6: 0 = Yin, 1 = Yang)."""
        out = torch.empty_like(norm_samples[:, 0])
        yang = norm_samples[:, -1] != 0
        for sel, vol, c0 in ((~yang, self.alpha_volume_yin, 0), (yang, self.alpha_volume_yang, 3)):
            pts = norm_samples[sel][:, c0:c0 + 3].view(1, -1, 1, 1, 3)          # x = r -> W, y = polar -> H, z = azimuth -> D
            out[sel] = F.grid_sample(vol, pts, align_corners=True).view(-1)
        return out


class _VolumeRender(torch.autograd.Function):
    """EgoNeRF.forward (EgoNeRF.py:491-602) as one autograd node around egn_render_forward / egn_render_backward."""

    @staticmethod
    def forward(ctx, model, opts, rays, u_coarse, u_fine, z_vals, *params):
        lib = _lib.load()
        n = rays.shape[0]
        ctx.dev_guard = _on(rays)
        with ctx.dev_guard:
            return _VolumeRender._forward(ctx, lib, n, model, opts, rays, u_coarse, u_fine, z_vals)

    @staticmethod
    def _forward(ctx, lib, n, model, opts, rays, u_coarse, u_fine, z_vals):
        tables = model._render_tables()          # (re)packs fp32 / bf16 tables if the parameters changed; before _config
        cfg = model._config(opts)
        S = lib.egn_samples_per_ray(cfg)
        has_env = cfg.env_h > 0
        dev = rays.device
        need_grad = bool(opts["is_train"]) and any(ctx.needs_input_grad[6:])
        rgb = torch.empty(n, 3, device=dev)
        depth = torch.empty(n, device=dev)
        alpha = torch.empty(n, S + (1 if has_env else 0), device=dev)
        bg = torch.empty(n, 3, device=dev) if has_env else None
        env = torch.empty(n, 3, device=dev) if has_env else None
        nbytes = lib.egn_workspace_bytes(cfg, n) if need_grad else lib.egn_workspace_bytes_eval(cfg, n)
        ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)
        out = _lib.EgnOutputs(rgb.data_ptr(), depth.data_ptr(), bg.data_ptr() if has_env else None,
                              env.data_ptr() if has_env else None, alpha.data_ptr())
        P = model._params_struct()
        if z_vals is not None:
            if tuple(z_vals.shape) != (n, S):
                raise ValueError(f"z_vals must be ({n}, {S})")
            _lib.check(lib.egn_render_samples(cfg, P, tables.data_ptr(), rays.data_ptr(), n, z_vals.data_ptr(), out,
                                              ws.data_ptr(), int(need_grad), _stream()))
        else:
            _lib.check(lib.egn_render_forward(cfg, P, tables.data_ptr(), rays.data_ptr(), n, int(opts["is_train"]),
                                              _lib.ptr(u_coarse), _lib.ptr(u_fine), int(opts["seed"]),
                                              int(opts["ray_index0"]), out, ws.data_ptr(), int(need_grad), _stream()))
        ctx.mark_non_differentiable(depth)
        if need_grad:
            ctx.model, ctx.opts, ctx.n = model, opts, n
            ctx.save_for_backward(rays, ws, tables)
        ctx.has_env = has_env
        ctx.need_grad = need_grad
        if has_env:
            return rgb, depth, bg, env, alpha
        return rgb, depth, alpha

    @staticmethod
    def backward(ctx, *douts):
        if not ctx.need_grad:
            raise RuntimeError("egonerf_b200: backward through a forward that was run without gradients")
        lib = _lib.load()
        model, opts = ctx.model, ctx.opts
        rays, ws, tables = ctx.saved_tensors
        with _on(rays):
            return _VolumeRender._backward(ctx, lib, model, opts, rays, ws, tables, douts)

    @staticmethod
    def _backward(ctx, lib, model, opts, rays, ws, tables, douts):
        if ctx.has_env:
            d_rgb, _, d_bg, d_env, d_alpha = douts
        else:
            d_rgb, _, d_alpha = douts
            d_bg = d_env = None
        c = lambda t: None if t is None else t.contiguous().float()
        d_rgb, d_bg, d_env, d_alpha = c(d_rgb), c(d_bg), c(d_env), c(d_alpha)
        cfg = model._config(opts)
        plist = model._param_list()
        table_opt = getattr(model, "_table_opt", None)
        # with a table-space optimiser the gradient may live in NVLink peer memory (TableAdam.enable_peer_exchange)
        d_tables = table_opt.grad_target(tables) if table_opt is not None else torch.zeros_like(tables)
        # one allocation for every gradient: the 24 factor gradients are overwritten by egn_unpack_table_grads, the rest
        # (basis, MLP, envmap) is accumulated into and must start at zero — a single fill instead of one per tensor.
        # With a table-space optimiser attached (egonerf_b200/optim.py) the factor gradients stay in table layout.
        # Non-factor gradients (basis, MLP, envmap) are ACCUMULATED by the kernels.  If such a parameter already carries a
        # contiguous fp32 .grad (e.g. the views of the exchange bucket that TableAdam.zero_grad() re-attaches and zeroes every
        # step), the kernels add straight into it and autograd gets None for it: no per-parameter copies, and several backward
        # passes per step sum as autograd would.
        def in_place(i, p):
            g = p.grad
            return (i >= 24 and g is not None and g.dtype == torch.float32 and g.is_contiguous() and g.device == tables.device
                    and g.shape == p.shape)
        inplace = [in_place(i, p) for i, p in enumerate(plist)]
        sizes = [0 if ((table_opt is not None and i < 24) or inplace[i]) else p.numel() for i, p in enumerate(plist)]
        flat = torch.empty(sum(sizes), device=tables.device, dtype=torch.float32)
        n_fac = sum(sizes[:24])
        if flat.numel() > n_fac:
            flat[n_fac:].zero_()
        grads, off = [], 0
        for i, (p, sz) in enumerate(zip(plist, sizes)):
            grads.append(p.grad if inplace[i] else (flat[off:off + sz].view_as(p) if sz else None))
            off += sz
        G = model._grads_struct(grads)
        # ray-sharded training exchanges the envmap gradient in sparse form (24 B per ray instead of the dense (3, 2h, h)
        # tensor): the kernel then writes d(loss)/d(env radiance) per ray and `allreduce_gradients` scatters all ranks' rays
        env_rays = None
        if ctx.has_env and getattr(model, "sparse_env_grad", False):
            env_rays = torch.empty(ctx.n, 3, device=tables.device, dtype=torch.float32)
            model._env_rays.append((rays[:, 3:6].contiguous(), env_rays))
        _lib.check(lib.egn_render_backward_sparse_env(cfg, model._params_struct(), tables.data_ptr(), rays.data_ptr(), ctx.n,
                                                      ws.data_ptr(), _lib.ptr(d_rgb), _lib.ptr(d_bg), _lib.ptr(d_env),
                                                      _lib.ptr(d_alpha), d_tables.data_ptr(), G, _lib.ptr(env_rays), _stream()))
        if env_rays is not None:
            grads[-1] = None                      # the dense emission gradient is produced by allreduce_gradients
        grads = [None if inplace[i] else g for i, g in enumerate(grads)]
        if table_opt is not None:
            table_opt.accumulate(d_tables)
        else:
            _lib.check(lib.egn_unpack_table_grads(cfg, d_tables.data_ptr(), G, _stream()))
        return (None, None, None, None, None, None) + tuple(grads)


class EgoNeRF(torch.nn.Module):
    """Yin-Yang VM-decomposed radiance field (reference: models/EgoNeRF.py:27, ctor kwargs models/tensorBase.py:133-139)."""

    def __init__(self, aabb, gridSize, device, coordinates, density_n_comp=8, appearance_n_comp=24, app_dim=27,
                 shadingMode='MLP_PE', alphaMask=None, near_far=[2.0, 6.0], density_shift=-10, alphaMask_thres=0.001,
                 distance_scale=25, rayMarch_weight_thres=0.0001, pos_pe=6, view_pe=6, fea_pe=6, featureC=128,
                 step_ratio=2.0, fea2denseAct='softplus', use_envmap=False, envmap_res_H=1000, envmap=None,
                 coarse_sigma_grid_update_rule=None, coarse_sigma_grid_reso=None, interval_th=False):
        super().__init__()
        if not isinstance(coordinates, YinYangSphericalCoords):
            raise TypeError("EgoNeRF needs egonerf_b200.models.coordinates.YinYangSphericalCoords")
        if isinstance(density_n_comp, int):
            density_n_comp = [density_n_comp] * 3
        if isinstance(appearance_n_comp, int):
            appearance_n_comp = [appearance_n_comp] * 3
        self.density_n_comp, self.app_n_comp, self.app_dim = list(density_n_comp), list(appearance_n_comp), app_dim
        self.aabb = aabb
        self.alphaMask = alphaMask
        self.device = device
        self.density_shift, self.alphaMask_thres, self.distance_scale = density_shift, alphaMask_thres, distance_scale
        self.rayMarch_weight_thres, self.fea2denseAct = rayMarch_weight_thres, fea2denseAct
        self.near_far, self.step_ratio = near_far, step_ratio
        self.gridSize = torch.LongTensor(list(gridSize)).to(device)
        self.matMode_yin = self.matMode_yang = MAT_MODE
        self.vecMode_yin = self.vecMode_yang = VEC_MODE
        self.coordinates = coordinates
        self.coarse_sigma_grid_update_rule = coarse_sigma_grid_update_rule
        self.shadingMode, self.pos_pe, self.view_pe, self.fea_pe, self.featureC = shadingMode, pos_pe, view_pe, fea_pe, featureC

        self.envmap = None
        if use_envmap:
            if envmap is None:
                self.init_envmap(envmap_res_H, init_strategy='random', device=device)
            else:
                self.envmap = EnvironmentMap(h=envmap.emission.shape[2], init_strategy='zero', device=device)
                self.envmap.load_envmap(envmap.emission, device=device)
        self._check_envelope()
        self.init_render_func(shadingMode, pos_pe, view_pe, fea_pe, featureC, device)
        self.init_svd_volume(gridSize[0], device)
        self.update_stepSize(list(gridSize))
        self._tables = None
        self._tables_key = None
        self._sched = {}
        self._cfg_static = None
        # arithmetic inside libegn_b200.  "tc_f16" (default): ONE fused tcgen05 kernel for the fine pass -- fp32 density /
        # alpha / compositing, fp16 appearance tables and MMA operands with fp32 accumulation: rgb within 2e-5 of the
        # reference on every fixture (bound 1e-4) -- and the tcgen05 backward (fp16 operands, gradients within 1e-2 of the
        # reference's).  "tc_split": unfused kernels, tcgen05 MLP with a 3-term bf16 split (fp32-equivalent, rgb ~1e-6) and
        # the exact fp32 backward (gradients 5e-4) -- 1.7x / 4x slower.  "fp32": FFMA everywhere.  "tc_bf16" is the
        # pre-ABI-7 name of "tc_f16".  Not a reference kwarg: set the attribute or EGN_MLP_MODE.
        self.mlp_mode = os.environ.get("EGN_MLP_MODE", "tc_f16")
        # "bf16": the tcgen05 BACKWARD kernels re-gather from a bf16 copy of the render tables (half the bytes); the
        # throughput-mode forward always reads the half tables (`_tables_h`: fp16 appearance + fp32 density)
        self.table_dtype = os.environ.get("EGN_TABLE_DTYPE", "f32")
        # True: backward on the tcgen05 kernels (bf16 operands) even when the forward runs in a parity mode
        self.tc_backward = os.environ.get("EGN_TC_BACKWARD", "0") == "1"
        self._tables_bf16 = None
        self._tables_h = None
        # ray-sharded training only (set by the multi-GPU launcher): keep the envmap gradient per ray until
        # `allreduce_gradients` has gathered every rank's rays (see `_VolumeRender.backward`)
        self.sparse_env_grad = False
        self._env_rays = []

    def _check_envelope(self):
        """What libegn_b200 is built for (egn_abi.cu `validate`; INTEGRATION.md "Supported envelope"): fail at construction
        with the reason, not at the first forward."""
        if self.density_n_comp != [16] * 3 or self.app_n_comp != [48] * 3:
            raise NotImplementedError(f"libegn_b200 is built for n_lamb_sigma=[16,16,16], n_lamb_sh=[48,48,48] (every shipped "
                                      f"EgoNeRF config); got {self.density_n_comp}, {self.app_n_comp}")
        if not 1 <= self.app_dim <= 27:
            raise NotImplementedError(f"app_dim={self.app_dim}: libegn_b200 supports 1..27")
        if self.shadingMode in ('MLP_Fea', 'MLP'):
            if self.featureC != 128:
                raise NotImplementedError(f"featureC={self.featureC}: libegn_b200 supports 128")
            in_dim = self.app_dim + 3 + 6 * self.view_pe + (2 * self.fea_pe * self.app_dim if self.shadingMode == 'MLP_Fea' else 0)
            if in_dim > 152:
                raise NotImplementedError(f"MLP input width {in_dim} (view_pe={self.view_pe}, fea_pe={self.fea_pe}) exceeds the 152 "
                                          "libegn_b200 supports; the shipped configs use view_pe = fea_pe = 2 (150)")

    # ---- parameters (EgoNeRF.py:96-122) -------------------------------------------------------------
    def init_render_func(self, shadingMode, pos_pe, view_pe, fea_pe, featureC, device):
        if shadingMode == 'MLP_Fea':
            self.renderModule = MLPRender_Fea(self.app_dim, view_pe, fea_pe, featureC).to(device)
        elif shadingMode == 'MLP':
            self.renderModule = MLPRender(self.app_dim, view_pe, featureC).to(device)
        elif shadingMode == 'SH':
            self.renderModule = SHRender
        elif shadingMode == 'RGB':
            assert self.app_dim == 3
            self.renderModule = RGBRender
        else:
            # MLP_PE crashes inside the reference's own EgoNeRF.forward (7-D points)
            raise NotImplementedError(f"shadingMode {shadingMode!r} is not supported by the EgoNeRF path")

    def init_one_svd(self, n_component, gridSize, scale, device):
        planes, lines = {}, {}
        for h in ('yin', 'yang'):
            planes[h] = torch.nn.ParameterList([torch.nn.Parameter(scale * torch.randn(
                (1, n_component[i], gridSize[MAT_MODE[i][1]], gridSize[MAT_MODE[i][0]]))) for i in range(3)]).to(device)
            lines[h] = torch.nn.ParameterList([torch.nn.Parameter(scale * torch.randn(
                (1, n_component[i], gridSize[VEC_MODE[i]], 1))) for i in range(3)]).to(device)
        return planes['yin'], lines['yin'], planes['yang'], lines['yang']

    def init_svd_volume(self, res, device):
        g = self.gridSize.tolist()
        self.density_plane_yin, self.density_line_yin, self.density_plane_yang, self.density_line_yang = \
            self.init_one_svd(self.density_n_comp, g, 0.1, device)
        self.app_plane_yin, self.app_line_yin, self.app_plane_yang, self.app_line_yang = \
            self.init_one_svd(self.app_n_comp, g, 0.1, device)
        self.basis_mat_yin = torch.nn.Linear(sum(self.app_n_comp), self.app_dim, bias=False).to(device)
        self.basis_mat_yang = torch.nn.Linear(sum(self.app_n_comp), self.app_dim, bias=False).to(device)

    def update_stepSize(self, gridSize):
        """TensorBase.update_stepSize (tensorBase.py:206-215): step of the uniform march (`exp_sampling=False`)."""
        aabb = torch.as_tensor(self.aabb).detach().float().cpu()
        self.aabbSize = aabb[1] - aabb[0]
        self.units = self.aabbSize / (torch.LongTensor(list(gridSize)) - 1)
        self.stepSize = torch.mean(self.units) * self.step_ratio
        self.aabbHalfDiag = torch.sqrt(torch.sum(torch.square(self.aabbSize))) / 2.0
        self.nSamples = int((self.aabbHalfDiag / self.stepSize).item()) + 1

    def init_envmap(self, envmap_res_H, init_strategy='zero', device='cuda'):
        self.envmap = EnvironmentMap(h=envmap_res_H, init_strategy=init_strategy, device=device)

    def get_optparam_groups(self, lr_init_spatialxyz=0.02, lr_init_network=0.001, lr_init_envmap=0.1, merged=False):
        """EgoNeRF.py:139-156.  `merged=True` (extension) returns the same parameters with the same learning rates in three
        groups instead of twelve, so that a fused multi-tensor optimiser needs three launches per step instead of twelve."""
        groups = []
        for h in ('yin', 'yang'):
            groups += [{'params': getattr(self, f'density_line_{h}'), 'lr': lr_init_spatialxyz},
                       {'params': getattr(self, f'density_plane_{h}'), 'lr': lr_init_spatialxyz},
                       {'params': getattr(self, f'app_line_{h}'), 'lr': lr_init_spatialxyz},
                       {'params': getattr(self, f'app_plane_{h}'), 'lr': lr_init_spatialxyz},
                       {'params': getattr(self, f'basis_mat_{h}').parameters(), 'lr': lr_init_network}]
        if isinstance(self.renderModule, torch.nn.Module):
            groups += [{'params': self.renderModule.parameters(), 'lr': lr_init_network}]
        if self.envmap is not None:
            groups += [{'params': self.envmap.emission, 'lr': lr_init_envmap}]
        if merged:
            by_lr = {}
            for g in groups:
                ps = [g['params']] if torch.is_tensor(g['params']) else list(g['params'])
                by_lr.setdefault(g['lr'], []).extend(ps)
            groups = [{'params': ps, 'lr': lr} for lr, ps in by_lr.items()]
        return groups

    # ---- checkpoint surface (tensorBase.py:241-268, EgoNeRF.py:158-187) -------------------------------
    def get_kwargs(self):
        return {'aabb': self.aabb, 'gridSize': self.gridSize.tolist(), 'density_n_comp': self.density_n_comp,
                'appearance_n_comp': self.app_n_comp, 'app_dim': self.app_dim, 'density_shift': self.density_shift,
                'alphaMask_thres': self.alphaMask_thres, 'distance_scale': self.distance_scale,
                'rayMarch_weight_thres': self.rayMarch_weight_thres, 'fea2denseAct': self.fea2denseAct,
                'near_far': self.near_far, 'step_ratio': self.step_ratio, 'shadingMode': self.shadingMode,
                'pos_pe': self.pos_pe, 'view_pe': self.view_pe, 'fea_pe': self.fea_pe, 'featureC': self.featureC,
                'coordinates': self.coordinates, 'use_envmap': self.envmap is not None, 'envmap': self.envmap,
                'coarse_sigma_grid_update_rule': self.coarse_sigma_grid_update_rule}

    def save(self, path, global_step):
        ckpt = {'kwargs': self.get_kwargs(), 'state_dict': self.state_dict(), 'global_step': global_step}
        if self.alphaMask is not None:                                   # bit-packed, EgoNeRF.py:161-167
            for h in ('yin', 'yang'):
                vol = getattr(self.alphaMask, f'alpha_volume_{h}').bool().cpu().numpy()
                ckpt.update({f'alphaMask_{h}.shape': vol.shape, f'alphaMask_{h}.mask': np.packbits(vol.reshape(-1))})
        if self.envmap is not None:
            ckpt.update({'envmap.emission': self.envmap.emission.detach().cpu().numpy(),
                         'envmap_res_H': self.envmap.emission.shape[2]})
        torch.save(ckpt, path)

    def load(self, ckpt):
        if 'alphaMask_yin.shape' in ckpt.keys():                         # EgoNeRF.py:175-180
            vols = []
            for h in ('yin', 'yang'):
                shape = ckpt[f'alphaMask_{h}.shape']
                bits = np.unpackbits(ckpt[f'alphaMask_{h}.mask'])[:int(np.prod(shape))].reshape(shape)
                vols.append(torch.from_numpy(bits).float().to(self.device))
            self.alphaMask = YinYangAlphaGridMask(self.device, *vols)
        if self.envmap is not None:
            self.envmap = EnvironmentMap(h=ckpt['envmap_res_H'], init_strategy='zero', device=self.device)
            self.envmap.load_envmap(emission=ckpt['envmap.emission'], device=self.device)
        self.load_state_dict(ckpt['state_dict'])
        self._fp_cache = self._pl_cache = self._ps_cache = None
        self.update_coarse_sigma_grid()
        return ckpt['global_step']

    # ---- render tables ------------------------------------------------------------------------------
    def _factor_params(self):
        """The 24 factor tensors [h][kind][i].  Walking eight ParameterLists costs ~90 us per call and a forward needs the list
        three times, so it is cached; the cache is dropped wherever Parameters are replaced by new objects
        (`upsample_volume_grid`, `load`) and re-validated against the first and the last entry on every use."""
        c = getattr(self, "_fp_cache", None)
        if c is not None and c[0] is self.density_plane_yin[0] and c[-1] is self.app_line_yang[2]:
            return c
        out = []
        for h in ('yin', 'yang'):
            for kind in ('density_plane', 'density_line', 'app_plane', 'app_line'):
                out += list(getattr(self, f'{kind}_{h}'))
        self._fp_cache = out
        self._pl_cache = None
        return out

    def _param_list(self):
        fp = self._factor_params()
        c = getattr(self, "_pl_cache", None)
        n_expected = 26 + (6 if isinstance(self.renderModule, torch.nn.Module) else 0) + (self.envmap is not None)
        if c is not None and len(c) == n_expected and (self.envmap is None or c[-1] is self.envmap.emission) \
                and c[24] is self.basis_mat_yin.weight:
            return c
        ps = list(fp) + [self.basis_mat_yin.weight, self.basis_mat_yang.weight]
        if isinstance(self.renderModule, torch.nn.Module):
            ps += [self.renderModule.mlp[0].weight, self.renderModule.mlp[0].bias, self.renderModule.mlp[2].weight,
                   self.renderModule.mlp[2].bias, self.renderModule.mlp[4].weight, self.renderModule.mlp[4].bias]
        if self.envmap is not None:
            ps += [self.envmap.emission]
        self._pl_cache = ps
        return ps

    def _fill_struct(self, S, tensors):
        f = tensors[:24]
        for hi in range(2):
            for ki, kind in enumerate(('density_plane', 'density_line', 'app_plane', 'app_line')):
                for i in range(3):
                    getattr(S, kind)[hi][i] = _lib.ptr(f[hi * 12 + ki * 3 + i])
        S.basis[0], S.basis[1] = _lib.ptr(tensors[24]), _lib.ptr(tensors[25])
        k = 26
        if isinstance(self.renderModule, torch.nn.Module):
            for l in range(3):
                S.mlp_w[l] = _lib.ptr(tensors[k + 2 * l])
                S.mlp_b[l] = _lib.ptr(tensors[k + 2 * l + 1])
            k += 6
        if self.envmap is not None:
            S.emission = _lib.ptr(tensors[k])
        return S

    def _params_struct(self):
        """EgnParams with the raw pointers of the current parameters; rebuilt only when a pointer changed."""
        plist = self._param_list()
        key = tuple(map(torch.Tensor.data_ptr, plist))
        c = getattr(self, "_ps_cache", None)
        if c is None or c[0] != key:
            c = (key, self._fill_struct(_lib.EgnParams(), [p.detach() for p in plist]))
            self._ps_cache = c
        return c[1]

    def _grads_struct(self, grads):
        return self._fill_struct(_lib.EgnGrads(), grads)

    def update_coarse_sigma_grid(self):
        """Reference: AvgPool refresh of the coarse density grid (EgoNeRF.py:124-133, called every iteration by
        train.py:356-357).  Here: re-pack the render tables (interleaved fine tables + pooled coarse tables).  With a
        table-space optimiser attached the tables are already up to date after its step (it writes them itself)."""
        if getattr(self, "_table_opt", None) is not None:
            return          # its step wrote parameters AND tables; any other parameter change moves the (data_ptr, version) key
        self._tables_key = None

    def _render_tables(self):
        fp = self._factor_params()
        _need_cuda(fp[0], "model parameters")
        with _on(fp[0]):
            return self._render_tables_on_device(fp)

    def _render_tables_on_device(self, fp):
        key = tuple((p.data_ptr(), p._version) for p in fp)
        if self._tables is None or key != self._tables_key:
            lib = _lib.load()
            cfg = self._config(None)
            nfl = lib.egn_table_floats(cfg)
            if nfl < 0:
                _lib.check(1)
            if self._tables is None or self._tables.numel() != nfl:
                self._tables = torch.empty(int(nfl), device=fp[0].device, dtype=torch.float32)
            else:
                self._tables = torch.empty_like(self._tables)     # saved-for-backward tables must not be overwritten
            _lib.check(lib.egn_pack_tables(cfg, self._params_struct(), self._tables.data_ptr(), _stream()))
            self._tables_key = key
            self._tables_bf16 = None
            self._tables_h = None
        if self._fused_mode() and getattr(self, "_tables_h", None) is None:
            lib = _lib.load()
            cfg = self._config(None)
            self._tables_h = torch.empty(int(lib.egn_table_h_bytes(cfg)), device=fp[0].device, dtype=torch.uint8)
            _lib.check(lib.egn_pack_tables_h(cfg, self._tables.data_ptr(), self._tables_h.data_ptr(), _stream()))
        if self.table_dtype in ("bf16", "f16") and self._fused_mode() and self._tables_bf16 is None:
            lib = _lib.load()
            cfg = self._config(None)
            ne = int(lib.egn_table_bf16_elems(cfg))
            self._tables_bf16 = torch.empty(ne, device=fp[0].device, dtype=torch.bfloat16)
            _lib.check(lib.egn_pack_tables_bf16(cfg, self._tables.data_ptr(), self._tables_bf16.data_ptr(), _stream()))
        return self._tables

    def _fused_mode(self):
        return self.mlp_mode in ("tc_f16", "tc_bf16")

    def _static_config(self):
        """Scalars that never change after construction, read back from the device ONCE (a `.cpu()` per forward would be
        a host-device synchronisation per call)."""
        if getattr(self, "_cfg_static", None) is None:
            co = self.coordinates
            if len(set(self.density_n_comp)) != 1 or len(set(self.app_n_comp)) != 1:
                raise NotImplementedError("n_lamb_sigma / n_lamb_sh must be equal across the three factor pairs")
            if [co.N_r, co.N_theta, co.N_phi] != self.gridSize.tolist():
                raise RuntimeError(f"factor grid {self.gridSize.tolist()} != coordinate resolution "
                                   f"{[co.N_r, co.N_theta, co.N_phi]}: call coordinates.set_resolution after "
                                   "upsample_volume_grid (train.py:376-377)")
            near, inv = co.near.cpu(), co.inv_diff.cpu()
            self._cfg_static = dict(
                grid=self.gridSize.tolist(), center=co.center.cpu().tolist(),
                aabb=torch.as_tensor(self.aabb).detach().float().cpu().reshape(-1).tolist(),
                ang_near=[float(near[1]), float(near[2])], ang_inv=[float(inv[1]), float(inv[2])],
                step_size=float(self.stepSize))
        return self._cfg_static

    def _config(self, opts):
        """EgnConfig for `opts`, cached: filling ~45 ctypes fields costs ~50 us per forward otherwise.  The key holds everything
        the struct is built from that can change after construction."""
        co = self.coordinates
        t16, th = self._tables_bf16, getattr(self, "_tables_h", None)
        key = (None if opts is None else (opts["n_coarse"], opts["n_fine"], opts["use_coarse_sample"], opts["resampling"],
                                          opts.get("exp_sampling", True)),
               self.mlp_mode, self.table_dtype, bool(self.tc_backward), None if t16 is None else t16.data_ptr(),
               None if th is None else th.data_ptr(), None if self.envmap is None else self.envmap.emission.shape[2],
               co.N_r, co.r0, co.interval_th, self.shadingMode, self.app_dim, self.view_pe, self.fea_pe, self.featureC,
               self.fea2denseAct, self.density_shift, self.distance_scale, self.near_far[0], self.near_far[1],
               self._cfg_static is None)
        cache = self.__dict__.setdefault("_cfg_cache", {})
        cfg = cache.get(key)
        if cfg is None:
            if len(cache) > 64:
                cache.clear()
            cfg = self._build_config(opts)
            key = key[:-1] + (self._cfg_static is None,)          # _build_config fills the static part
            cache[key] = cfg
        return cfg

    def _build_config(self, opts):
        co = self.coordinates
        if self._sched.get('ladder') != (co.N_r, co.r0):          # set_resolution changes N_r and resets r0
            self._sched = {'ladder': (co.N_r, co.r0)}
            self._cfg_static = None
        st = self._static_config()
        cfg = _lib.EgnConfig()
        cfg.grid[:] = st["grid"]
        cfg.c_sigma, cfg.c_app, cfg.app_dim = self.density_n_comp[0], self.app_n_comp[0], self.app_dim
        cfg.shading = _lib.SHADING[self.shadingMode]
        cfg.view_pe, cfg.fea_pe, cfg.feature_c = self.view_pe, self.fea_pe, self.featureC
        cfg.fea2dense = _lib.ACT[self.fea2denseAct]
        tc_ok = self.shadingMode == 'MLP_Fea' and self.view_pe == 2 and self.fea_pe == 2
        cfg.mlp_mode = _lib.MLP_MODE[self.mlp_mode] if tc_ok else 0
        cfg.bwd_tc = int(bool(self.tc_backward) and tc_ok)
        use_bf16 = self.table_dtype in ("bf16", "f16") and self._fused_mode() and self._tables_bf16 is not None
        cfg.tables_bf16 = self._tables_bf16.data_ptr() if use_bf16 else None
        use_h = self._fused_mode() and tc_ok and getattr(self, "_tables_h", None) is not None
        cfg.tables_h = self._tables_h.data_ptr() if use_h else None
        cfg.env_h = self.envmap.emission.shape[2] if self.envmap is not None else 0
        cfg.center[:] = st["center"]
        cfg.near_plane = self.near_far[0]
        cfg.far_plane = self.near_far[1]
        cfg.step_size = st["step_size"]
        cfg.aabb[:] = st["aabb"]
        cfg.exp_sampling = 1
        cfg.density_shift, cfg.distance_scale = self.density_shift, self.distance_scale
        cfg.ang_near[:] = st["ang_near"]
        cfg.ang_inv[:] = st["ang_inv"]
        dev = self.density_plane_yin[0].device
        if 'knots' not in self._sched:
            self._sched['knots'] = co.r_knots().to(dev).contiguous()
            if not co.interval_th:                                # coarse pass: normalize_coord(downsample=2), EgoNeRF.py:523
                self._sched['knots_coarse'] = co.r_knots(downsample=2).to(dev).contiguous()
        cfg.r_knots = self._sched['knots'].data_ptr()
        cfg.plain_ladders = int(not co.interval_th)
        if not co.interval_th:
            cfg.r_knots_coarse = self._sched['knots_coarse'].data_ptr()
        if opts is not None:
            nc = int(opts["n_coarse"])
            cfg.n_coarse, cfg.n_fine = nc, int(opts["n_fine"])
            cfg.use_coarse_sample, cfg.resampling = int(opts["use_coarse_sample"]), int(opts["resampling"])
            cfg.exp_sampling = int(opts.get("exp_sampling", True))
            if ('z', nc) not in self._sched:
                if co.interval_th:
                    self._sched[('z', nc)] = sample_schedule(self.near_far[0], self.near_far[1], co.r0, nc).to(dev).contiguous()
                else:
                    r, ratio, r0p = plain_sample_schedule(self.near_far[0], self.near_far[1], nc)
                    self._sched[('z', nc)] = r.to(dev).contiguous()
                    self._sched[('jitter', nc)] = (ratio, r0p)
            cfg.z_coarse = self._sched[('z', nc)].data_ptr()
            if not co.interval_th:
                cfg.jitter_ratio, cfg.jitter_r0 = self._sched[('jitter', nc)]
        return cfg

    # ---- stand-alone operators ------------------------------------------------------------------------
    def feature2density(self, density_features):
        """tensorBase.py:415-419 (element-wise; plain torch)."""
        if self.fea2denseAct == "softplus":
            return F.softplus(density_features + self.density_shift)
        return F.relu(density_features)

    def sample_depths(self, rays_chunk, is_train=False, n_coarse=128, n_fine=128, resampling=True, use_coarse_sample=True,
                      u_coarse=None, u_fine=None, seed=0, ray_index0=0, exp_sampling=True):
        """Sorted sample depths (N,S): sample_ray_exp + coarse pass + sample_pdf + sort (EgoNeRF.py:507-542)."""
        _need_cuda(rays_chunk, "rays_chunk")
        lib = _lib.load()
        opts = dict(is_train=bool(is_train), n_coarse=n_coarse, n_fine=n_fine if resampling else 0,
                    resampling=bool(resampling), use_coarse_sample=bool(use_coarse_sample), exp_sampling=bool(exp_sampling))
        cfg = self._config(opts)
        rays = rays_chunk.detach().contiguous().float()
        n, S = rays.shape[0], lib.egn_samples_per_ray(cfg)
        z = torch.empty(n, S, device=rays.device)
        uc = u_coarse.contiguous().float() if u_coarse is not None else None
        uf = u_fine.contiguous().float() if u_fine is not None else None
        with _on(rays):
            _lib.check(lib.egn_sample_rays(cfg, self._render_tables().data_ptr(), rays.data_ptr(), n, int(is_train),
                                           _lib.ptr(uc), _lib.ptr(uf), int(seed), int(ray_index0), z.data_ptr(), _stream()))
        return z

    def _gather(self, coords_sampled, coarse=False, want_app=False):
        _need_cuda(coords_sampled, "coords_sampled")
        lib = _lib.load()
        c7 = coords_sampled.detach().reshape(-1, 7).contiguous().float()
        m = c7.shape[0]
        cfg = self._config(None)
        sig = torch.empty(m, device=c7.device)
        with _on(c7):
            tables = self._render_tables()
            if want_app:
                feat = torch.empty(m, 28, device=c7.device)
                _lib.check(lib.egn_app_feature(cfg, self._params_struct(), tables.data_ptr(), c7.data_ptr(), m,
                                               sig.data_ptr(), feat.data_ptr(), _stream()))
                return feat[:, :self.app_dim].reshape(*coords_sampled.shape[:-1], self.app_dim)
            _lib.check(lib.egn_density_feature(cfg, tables.data_ptr(), c7.data_ptr(), m, int(coarse), sig.data_ptr(), _stream()))
        return sig.view(coords_sampled.shape[:-1])

    def compute_densityfeature(self, coords_sampled):
        """EgoNeRF.py:291-347 (forward only; gradients flow through `forward`)."""
        return self._gather(coords_sampled)

    def compute_coarse_densityfeature(self, coords_sampled, coarse_sigma_grid_update_rule='conv'):
        """EgoNeRF.py:232-289."""
        return self._gather(coords_sampled, coarse=True)

    def compute_appfeature(self, coords_sampled):
        """EgoNeRF.py:349-413."""
        return self._gather(coords_sampled, want_app=True)

    # ---- regularisers on the factor tensors: plain torch on the same Parameters ------
    def vectorDiffs(self, vector_comps):
        total = 0
        for v in vector_comps:
            n_comp, n_size = v.shape[1:-1]
            dotp = torch.matmul(v.view(n_comp, n_size), v.view(n_comp, n_size).transpose(-1, -2))
            total = total + torch.mean(torch.abs(dotp.view(-1)[1:].view(n_comp - 1, n_comp + 1)[..., :-1]))
        return total

    def vector_comp_diffs(self):
        return sum(self.vectorDiffs(getattr(self, f'{k}_line_{h}')) for k in ('density', 'app') for h in ('yin', 'yang'))

    def density_L1(self):
        total = 0
        for h in ('yin', 'yang'):
            for i in range(3):
                total = total + torch.mean(torch.abs(getattr(self, f'density_plane_{h}')[i])) \
                    + torch.mean(torch.abs(getattr(self, f'density_line_{h}')[i]))
        return total

    def TV_loss_density(self, reg):
        return sum(reg(getattr(self, f'density_plane_{h}')[i]) * 1e-2 for i in range(3) for h in ('yin', 'yang'))

    def TV_loss_app(self, reg):
        return sum(reg(getattr(self, f'app_plane_{h}')[i]) * 1e-2 for i in range(3) for h in ('yin', 'yang'))

    @torch.no_grad()
    def up_sampling_VM(self, plane_coef, line_coef, res_target):
        """EgoNeRF.up_sampling_VM (EgoNeRF.py:415-425): planes are (1, C, G[m1], G[m0]) -> ids [m1, m0]; lines ids [v]."""
        for i in range(3):
            m0, m1 = MAT_MODE[i]
            plane_coef[i] = self.coordinates.up_sampling_VM(plane_coef[i].data, res_target=res_target, ids=[m1, m0])
            line_coef[i] = self.coordinates.up_sampling_VM(line_coef[i].data, res_target=res_target, ids=[VEC_MODE[i]])
        return plane_coef, line_coef

    @torch.no_grad()
    def upsample_volume_grid(self, res_target):
        """EgoNeRF.upsample_volume_grid (EgoNeRF.py:427-436; train.py:371-377): resample all 24 factor tensors to
        `res_target` = [N_r, N_theta, N_phi].  As in the reference the caller then calls
        `coordinates.set_resolution(res_target)` (which also resets r0 to 0.05, coordinates.py:206-215) and rebuilds the
        optimiser; rendering between the two calls is refused (`_static_config`)."""
        res_target = [int(v) for v in res_target]
        self.app_plane_yin, self.app_line_yin = self.up_sampling_VM(self.app_plane_yin, self.app_line_yin, res_target)
        self.density_plane_yin, self.density_line_yin = self.up_sampling_VM(self.density_plane_yin, self.density_line_yin, res_target)
        self.app_plane_yang, self.app_line_yang = self.up_sampling_VM(self.app_plane_yang, self.app_line_yang, res_target)
        self.density_plane_yang, self.density_line_yang = self.up_sampling_VM(self.density_plane_yang, self.density_line_yang, res_target)
        self.gridSize = torch.LongTensor(res_target).to(self.gridSize.device)
        self.update_stepSize(res_target)
        # everything derived from the old resolution: render tables, ladders, cached scalars, table-space optimiser state
        self._tables = self._tables_key = self._tables_bf16 = self._tables_h = self._cfg_static = None
        self._fp_cache = self._pl_cache = self._ps_cache = None
        self._cfg_cache = {}
        self._sched = {}
        self._table_opt = None
        self._bucket = None
        print(f'upsamping to {res_target}')

    def compute_alpha(self, norm_locs, length=1):
        """TensorBase.compute_alpha (tensorBase.py:421-436): alpha of one step of `length` at normalised 7-coords; samples
        the occupancy mask rejects get sigma = 0.  The density gather runs in libegn_b200 (`egn_density_feature`)."""
        if self.alphaMask is not None:
            alpha_mask = self.alphaMask.sample_alpha(norm_locs) > 0
        else:
            alpha_mask = torch.ones_like(norm_locs[:, 0], dtype=torch.bool)
        sigma = torch.zeros(norm_locs.shape[:-1], device=norm_locs.device)
        if alpha_mask.any():
            sigma[alpha_mask] = self.feature2density(self.compute_densityfeature(norm_locs[alpha_mask]))
        return 1 - torch.exp(-sigma * float(length)).view(norm_locs.shape[:-1])

    @torch.no_grad()
    def getDenseAlpha(self, gridSize=None):
        """EgoNeRF.getDenseAlpha (EgoNeRF.py:438-465): alpha of one march step on a regular lattice of normalised
        coordinates, for each hemisphere; (g0, g1, g2) each.  Evaluated in slabs of <= 2^22 lattice points."""
        gridSize = self.gridSize.tolist() if gridSize is None else [int(v) for v in gridSize]
        dev = self.density_plane_yin[0].device
        lin = [torch.linspace(0, 1, g, device=dev) * 2 - 1 for g in gridSize]
        alpha = [torch.empty(gridSize, device=dev), torch.empty(gridSize, device=dev)]
        slab = max(1, (1 << 22) // (gridSize[1] * gridSize[2]))
        for i0 in range(0, gridSize[0], slab):
            pts = torch.stack(torch.meshgrid(lin[0][i0:i0 + slab], lin[1], lin[2], indexing='ij'), -1).reshape(-1, 3)
            for h in (0, 1):
                c7 = torch.zeros(pts.shape[0], 7, device=dev)
                c7[:, 3 * h:3 * h + 3] = pts
                c7[:, 6] = h
                alpha[h][i0:i0 + slab] = self.compute_alpha(c7, self.stepSize).view(-1, gridSize[1], gridSize[2])
        return alpha[0], alpha[1]

    @torch.no_grad()
    def updateAlphaMask(self, gridSize=None):
        """EgoNeRF.updateAlphaMask (EgoNeRF.py:467-489): dense alpha -> 3x3x3 max-pool -> threshold at `alphaMask_thres` ->
        YinYangAlphaGridMask.  Deprecated in the reference (and unreachable from train.py:359-362); kept for callers of
        `compute_alpha`.  Returns None like the reference."""
        gridSize = self.gridSize.tolist() if gridSize is None else [int(v) for v in gridSize]
        vols = []
        for a in self.getDenseAlpha(gridSize):
            a = a.clamp(0, 1).transpose(0, 2).contiguous()[None, None]
            a = F.max_pool3d(a, kernel_size=3, padding=1, stride=1).view(gridSize[::-1])
            vols.append((a >= self.alphaMask_thres).float())
        self.alphaMask = YinYangAlphaGridMask(self.device, vols[0], vols[1])
        total = float(vols[0].sum() + vols[1].sum())
        print("alpha rest %%%f" % (total / (2 * gridSize[0] * gridSize[1] * gridSize[2]) * 100))

    def launches_per_forward(self, S=256, keep_for_backward=False):
        """Kernels of libegn_b200 launched by one `forward`: sampler, gather, [MLP], composite -- or, in the throughput mode,
        sampler + operand-image kernel + fused fine pass (+ composite only when the backward pass needs the per-sample state
        or S % 128 != 0)."""
        if not isinstance(self.renderModule, torch.nn.Module):
            return 3
        fused = self._fused_mode() and self.shadingMode == 'MLP_Fea' and self.view_pe == 2 and self.fea_pe == 2
        if fused:        # sampler + operand-image kernel (6 us) + fused fine pass (+ composite)
            return 3 if (not keep_for_backward and S % 128 == 0) else 4
        return 4

    def launches_per_train_step(self, n_rays):
        """forward + backward (composite, per-sub-chunk MLP chain, gather) + gradient unpack + table re-pack."""
        mlp = isinstance(self.renderModule, torch.nn.Module)
        sub = -(-int(n_rays) // 4096)
        return self.launches_per_forward(keep_for_backward=True) + 1 + (6 * sub if mlp else 0) + 1 + 1 + 1

    def allreduce_gradients(self, group=None, average=False):
        """Ray-sharded data parallelism: ONE all-reduce (sum) over all parameter gradients, through a
        persistent flat bucket (egonerf_b200/sharding.py)."""
        from ..sharding import GradientBucket, gather_env_gradient
        ps = self._param_list()
        table_opt = getattr(self, "_table_opt", None)
        if table_opt is not None:                                # factor gradients live in table layout: one buffer already
            ps = ps[24:]
        if self.envmap is not None and self.sparse_env_grad:     # 24 B / ray all-gather + local scatter instead of 88 MB
            gather_env_gradient(self, group, average)
            ps = ps[:-1]
        # peer-memory exchange (TableAdam.enable_peer_exchange): the bucket lives in the tail of the peer buffer and is
        # summed by the same kernel as the factor gradient
        peer_tail = getattr(table_opt, "peer_tail", None) if (table_opt is not None and table_opt.peer is not None) else None
        if peer_tail is not None and sum(p.numel() for p in ps) > peer_tail.numel():
            peer_tail = None
        if (getattr(self, "_bucket", None) is None or [id(p) for p in self._bucket.params] != [id(p) for p in ps]
                or (peer_tail is not None) != self._bucket.external):
            self._bucket = GradientBucket(ps, storage=peer_tail)
        self._bucket.gather_from_params()
        if table_opt is not None:
            table_opt.allreduce(group, average)
        if not self._bucket.external:
            self._bucket.allreduce(group, average)

    def stage_times(self, rays_chunk, repeats=3, n_coarse=128, n_fine=128, resampling=True, use_coarse_sample=True, **_):
        """Mean device time (ms) of each stage of the eval forward, from CUDA events recorded between the launches
        (egn_render_forward_timed)."""
        _need_cuda(rays_chunk, "rays_chunk")
        lib = _lib.load()
        opts = dict(is_train=False, n_coarse=n_coarse, n_fine=n_fine if resampling else 0, resampling=bool(resampling),
                    use_coarse_sample=bool(use_coarse_sample))
        tables = self._render_tables()
        cfg = self._config(opts)
        rays = rays_chunk.detach().contiguous().float()
        n, S, dev = rays.shape[0], lib.egn_samples_per_ray(cfg), rays.device
        has_env = cfg.env_h > 0
        rgb, depth = torch.empty(n, 3, device=dev), torch.empty(n, device=dev)
        alpha = torch.empty(n, S + (1 if has_env else 0), device=dev)
        bg = torch.empty(n, 3, device=dev) if has_env else None
        env = torch.empty(n, 3, device=dev) if has_env else None
        ws = torch.empty(int(lib.egn_workspace_bytes_eval(cfg, n)), dtype=torch.uint8, device=dev)
        out = _lib.EgnOutputs(rgb.data_ptr(), depth.data_ptr(), _lib.ptr(bg), _lib.ptr(env), alpha.data_ptr())
        ms = (C.c_float * 4)()
        tot = [0.0] * 4
        P = self._params_struct()
        for _ in range(repeats):
            _lib.check(lib.egn_render_forward_timed(cfg, P, tables.data_ptr(), rays.data_ptr(), n, 0, None, None, 0, 0,
                                                    out, ws.data_ptr(), _stream(), ms))
            tot = [a + b for a, b in zip(tot, ms)]
        return [t / repeats for t in tot]

    # ---- forward --------------------------------------------------------------------------------------
    def forward(self, rays_chunk, white_bg=True, is_train=False, ndc_ray=False, n_coarse=-1, n_fine=0,
                exp_sampling=False, pretrain_envmap=False, pivotal_sample_th=0., resampling=False,
                use_coarse_sample=True, interval_th=False, u_coarse=None, u_fine=None, seed=None, ray_index0=0,
                z_vals=None):
        """Same signature and return tuple as the reference (EgoNeRF.py:491-602).  Extra keyword-only extensions:
        `u_coarse` / `u_fine` inject the train-mode uniforms (reproducible parity tests), `seed` / `ray_index0` key
        the in-kernel generator so that ray-sharded runs draw disjoint streams, `z_vals` (N,S) supplies the sorted sample
        depths and skips the sampler (the reference's sampler / renderer split)."""
        _need_cuda(rays_chunk, "rays_chunk")
        if pretrain_envmap:
            return self.envmap.get_radiance(rays_chunk[:, 3:6])
        if ndc_ray:
            raise NotImplementedError          # EgoNeRF.py:503-504
        if n_coarse <= 0:
            raise ValueError("n_coarse must be given")
        rays = rays_chunk.detach().contiguous().float()
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if is_train and u_coarse is None else 0
        opts = dict(is_train=bool(is_train), n_coarse=n_coarse, n_fine=n_fine if resampling else 0,
                    resampling=bool(resampling), use_coarse_sample=bool(use_coarse_sample), seed=seed,
                    ray_index0=ray_index0, exp_sampling=bool(exp_sampling))
        uc = u_coarse.contiguous().float() if u_coarse is not None else None
        uf = u_fine.contiguous().float() if u_fine is not None else None
        zv = z_vals.detach().contiguous().float() if z_vals is not None else None
        outs = _VolumeRender.apply(self, opts, rays, uc, uf, zv, *self._param_list())
        if self.envmap is not None:
            rgb, depth, bg, env, alpha = outs
            return rgb, depth, bg, env, alpha
        rgb, depth, alpha = outs
        return rgb, depth, None, None, alpha
