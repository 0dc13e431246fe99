"""Shared test helpers: golden-case registry, oracle configuration, boundary-aware comparison."""
import os

import numpy as np
import torch

from egonerf_b200.synthetic import make_scene
from oracle import egn_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

TINY = dict(n_voxels=40 ** 3, seed=7)
TINY_ENV = dict(n_voxels=40 ** 3, seed=8, envmap_h=32, near_far=(0.1, 300.), r0=0.05, density_shift=-10.)
# fixture -> (make_scene kwargs, render kwargs overriding the defaults)
RENDER_CASES = {
    "render_tiny_eval": (TINY, {}),
    "render_tiny_train": (TINY, {}),
    "render_tiny_env_eval": (TINY_ENV, {}),
    "render_tiny_env_train_grad": (TINY_ENV, {}),
    "render_tiny_train_grad": (TINY, {}),
    "render_tiny_train_mse_grad": (TINY, {}),           # gradients of the MSE training loss (train.py:260)
    "render_tiny_env_train_mse_grad": (TINY_ENV, {}),
    "render_tiny_noresample": (TINY, dict(resampling=False, n_fine=0)),
    "render_tiny_fineonly": (TINY, dict(use_coarse_sample=False)),
    "render_128_eval": (dict(n_voxels=128 ** 3), {}),
    "render_128_white_eval": (dict(n_voxels=128 ** 3, smooth=1), {}),
    "render_300_eval": (dict(n_voxels=27e6), {}),
    "render_300_train": (dict(n_voxels=27e6), {}),
    "render_tiny_march_eval": (TINY, dict(exp_sampling=False)),
    "render_tiny_march_train": (TINY, dict(exp_sampling=False)),
    # a run without --interval_th: plain exponential ladders, coarse pass on the N_r/2 ladder
    "render_tiny_plain_eval": (TINY, dict(interval_th=False)),
    "render_tiny_plain_train_grad": (TINY, dict(interval_th=False)),
    "render_tiny_plain_noresample": (TINY, dict(interval_th=False, resampling=False, n_fine=0)),
    "render_tiny_mlp": (dict(n_voxels=40 ** 3, seed=9, shading='MLP'), {}),
    "render_tiny_rgb": (dict(n_voxels=40 ** 3, seed=10, shading='RGB', app_dim=3), {}),
}

_scene_cache = {}


def scene_for(kwargs):
    key = repr(sorted(kwargs.items()))
    if key not in _scene_cache:
        _scene_cache[key] = make_scene(**kwargs)
    return _scene_cache[key]


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def oracle_cfg(scene, **kw):
    return O.OracleCfg(aabb=scene.aabb, grid=tuple(scene.grid), r0=scene.r0, near=scene.near_far[0],
                       far=scene.near_far[1], density_shift=scene.density_shift,
                       distance_scale=scene.distance_scale, shading=scene.shading, view_pe=scene.view_pe,
                       fea_pe=scene.fea_pe, app_dim=scene.app_dim, **kw)


def checksum(sd):
    return np.array([float(v.double().sum()) for v in sd.values()] + [float(v.double().abs().sum()) for v in sd.values()])


def T(x):
    return torch.from_numpy(np.asarray(x))


def stable_rays(scene, cfg, rays, is_train, u_c, u_f, margin=2e-6):
    """Rays none of whose samples sits within `margin` rad of a Yin/Yang classification threshold
    (coordinates.py:483-486) or whose inverse-CDF denominators sit on the 1e-5 switch (ray_utils.py:183):
    there a 1-ulp difference in acosf/atan2f legitimately moves a sample to the other, independent grid."""
    with torch.no_grad():
        _, aux = O.render(scene.state_dict, cfg, rays, is_train, u_c, u_f, emission=scene.emission, want_aux=True)
    ok = aux["margin"] > margin
    if "cdf_den" in aux:
        ok &= ((aux["cdf_den"] - 1e-5).abs() > 1e-9).all(-1)
    return ok
