"""Deterministic synthetic scenes and ray batches: no dataset or checkpoint is reachable
offline, so tests and `bench.py` render random-init factor grids of the reference's exact shapes.

State-dict keys / shapes are the reference's (`EgoNeRF.init_one_svd`, EgoNeRF.py:102-122;
`MLPRender_Fea`, tensorBase.py:54-66) so the same dict loads into the reference model, into the oracle
and into the drop-in `EgoNeRF` module.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass
from typing import Optional

import torch

MAT_MODE = ((0, 1), (0, 2), (1, 2))
VEC_MODE = (2, 1, 0)


def yinyang_resolution(n_voxels: float):
    """Grid [N_r, N_theta, N_phi] per hemisphere (coordinates.py:507-520)."""
    n_r = int(pow(n_voxels, 1 / 3) / 2)
    n_t = int(n_r * 2 * math.sqrt(3) / 3)
    n_p = n_t * 3
    return [n_r + n_r % 2, n_t + n_t % 2, n_p + n_p % 2]


@dataclass
class Scene:
    aabb: torch.Tensor            # (2,3)
    grid: list                    # [N_r, N_theta, N_phi]
    near_far: list
    r0: float
    density_shift: float
    distance_scale: float
    n_lamb_sigma: list
    n_lamb_sh: list
    app_dim: int
    shading: str
    view_pe: int
    fea_pe: int
    featureC: int
    state_dict: "OrderedDict[str, torch.Tensor]"
    emission: Optional[torch.Tensor]      # (3,2h,h) or None
    n_voxels: float

    def model_kwargs(self):
        """kwargs of `EgoNeRF(aabb, gridSize, device, coordinates, **kw)` as train.py:163-171 passes them."""
        return dict(density_n_comp=self.n_lamb_sigma, appearance_n_comp=self.n_lamb_sh, app_dim=self.app_dim,
                    near_far=self.near_far, shadingMode=self.shading, alphaMask_thres=1e-4,
                    density_shift=self.density_shift, distance_scale=self.distance_scale, pos_pe=6,
                    view_pe=self.view_pe, fea_pe=self.fea_pe, featureC=self.featureC, step_ratio=0.5,
                    fea2denseAct='softplus', use_envmap=self.emission is not None,
                    envmap_res_H=(self.emission.shape[2] if self.emission is not None else 1000),
                    coarse_sigma_grid_update_rule='conv', coarse_sigma_grid_reso=None, interval_th=True)


def make_scene(n_voxels=128 ** 3, near_far=(0.01, 15.0), r0=0.03, density_shift=-8.0, distance_scale=25.0,
               sigma_std=0.7, app_std=0.1, n_lamb_sigma=(16, 16, 16), n_lamb_sh=(48, 48, 48), app_dim=27,
               shading='MLP_Fea', view_pe=2, fea_pe=2, featureC=128, envmap_h=None, traj_radius=0.5,
               seed=20221028, smooth=8) -> Scene:
    """aabb = ±(traj_radius + far) cube around the origin (dataset_omniblender.py:24-32).

    `smooth` = correlation length (texels) of the random factor fields: noise is drawn on a grid `smooth` times
    coarser and bilinearly upsampled, then rescaled to the requested std.  smooth=1 gives white noise — a stress
    case in which a 1e-4 shift of a sample depth already moves rgb by >1e-4 (the reference's own inverse-CDF
    resampling is that sensitive to 1-ulp differences in exp(), see tests/test_gpu_parity.py); trained fields
    are spatially coherent, which `smooth=8` imitates."""
    g = torch.Generator().manual_seed(seed)
    half = traj_radius + near_far[1]
    aabb = torch.tensor([[-half] * 3, [half] * 3], dtype=torch.float32)
    grid = yinyang_resolution(n_voxels)
    sd = OrderedDict()

    def rn(*shape, std):
        if smooth <= 1 or len(shape) != 4:
            return std * torch.randn(*shape, generator=g, dtype=torch.float32)
        _, c, hh, ww = shape
        lo = torch.randn(1, c, -(-hh // smooth) + 1, (-(-ww // smooth) + 1) if ww > 1 else 1, generator=g,
                         dtype=torch.float32)
        if ww == 1:
            lo = lo.expand(-1, -1, -1, 2)
            f = torch.nn.functional.interpolate(lo, size=(hh, 2), mode='bilinear', align_corners=True)[..., :1]
        else:
            f = torch.nn.functional.interpolate(lo, size=(hh, ww), mode='bilinear', align_corners=True)
        return (f * (std / f.std())).contiguous()

    def uni(*shape, bound):
        return (torch.rand(*shape, generator=g, dtype=torch.float32) * 2 - 1) * bound

    if shading in ('MLP_Fea', 'MLP'):
        in_c = app_dim + 3 + 2 * view_pe * 3 + (2 * fea_pe * app_dim if shading == 'MLP_Fea' else 0)
        for name, (fo, fi) in (('0', (featureC, in_c)), ('2', (featureC, featureC)), ('4', (3, featureC))):
            b = 1.0 / math.sqrt(fi)              # nn.Linear default init bound
            sd[f'renderModule.mlp.{name}.weight'] = uni(fo, fi, bound=b)
            sd[f'renderModule.mlp.{name}.bias'] = uni(fo, bound=b) if name != '4' else torch.zeros(fo)
    for hemi in ('yin', 'yang'):
        for kind, comps, std in (('density', n_lamb_sigma, sigma_std), ('app', n_lamb_sh, app_std)):
            for i in range(3):
                m0, m1 = MAT_MODE[i]
                sd[f'{kind}_plane_{hemi}.{i}'] = rn(1, comps[i], grid[m1], grid[m0], std=std)
            for i in range(3):
                sd[f'{kind}_line_{hemi}.{i}'] = rn(1, comps[i], grid[VEC_MODE[i]], 1, std=std)
        sd[f'basis_mat_{hemi}.weight'] = uni(app_dim, sum(n_lamb_sh), bound=1.0 / math.sqrt(sum(n_lamb_sh)))
    emission = None
    if envmap_h is not None:
        emission = torch.rand(3, 2 * envmap_h, envmap_h, generator=g, dtype=torch.float32) * 4 - 2
    return Scene(aabb=aabb, grid=grid, near_far=list(near_far), r0=r0, density_shift=density_shift,
                 distance_scale=distance_scale, n_lamb_sigma=list(n_lamb_sigma), n_lamb_sh=list(n_lamb_sh),
                 app_dim=app_dim, shading=shading, view_pe=view_pe, fea_pe=fea_pe, featureC=featureC,
                 state_dict=sd, emission=emission, n_voxels=n_voxels)


def make_rays(n, kind='isotropic', traj_radius=0.5, seed=1, erp_hw=None, row0=0) -> torch.Tensor:
    """(n,6) fp32 [origin, unit direction].
    isotropic : origins uniform in a horizontal disc of radius `traj_radius` (egocentric capture), dirs uniform on S2
    probe     : origins (U^3-.5)*.5 as in SURVEY §8c probes
    erp       : one pose at the origin, directions of an H x W equirect frame (ray_utils.py:24-40), rows from row0
    """
    g = torch.Generator().manual_seed(seed)
    if kind == 'erp':
        H, W = erp_hw
        rows = (n + W - 1) // W
        i = torch.arange(W, dtype=torch.float32)[None, :].expand(rows, W) + 0.5
        j = (torch.arange(rows, dtype=torch.float32) + row0)[:, None].expand(rows, W) + 0.5
        phi = (1 - 2 * i / W) * math.pi
        theta = (1 - 2 * j / H) * math.pi / 2
        d = torch.stack([-torch.cos(theta) * torch.sin(phi), torch.sin(theta),
                         -torch.cos(theta) * torch.cos(phi)], -1).reshape(-1, 3)[:n]
        d = d / torch.norm(d, dim=-1, keepdim=True)
        o = torch.zeros_like(d)
        return torch.cat([o, d], -1).contiguous()
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g, dtype=torch.float32), dim=-1)
    if kind == 'probe':
        o = (torch.rand(n, 3, generator=g, dtype=torch.float32) - .5) * .5
    else:
        rad = traj_radius * torch.sqrt(torch.rand(n, generator=g, dtype=torch.float32))
        ang = 2 * math.pi * torch.rand(n, generator=g, dtype=torch.float32)
        o = torch.stack([rad * torch.cos(ang), torch.zeros(n), rad * torch.sin(ang)], -1)
    return torch.cat([o, d], -1).contiguous()
