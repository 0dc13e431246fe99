"""Glue between `synthetic.Scene` and the drop-in module (tests, bench, smoke)."""
import torch

from .models.coordinates import YinYangSphericalCoords
from .models.EgoNeRF import EgoNeRF


def model_from_scene(scene, device="cuda", interval_th=True, mlp_mode="tc_split"):
    """Builds the drop-in EgoNeRF exactly as train.py:118-171 does and loads the scene's parameters.  The tests and the
    bench set the arithmetic mode explicitly; callers that do not get `mlp_mode` = "tc_split" here (fp32-equivalent forward,
    exact fp32 backward: what the 1e-6-level comparisons need), NOT the class default "tc_f16" a train.py user gets."""
    aabb = scene.aabb.to(device)
    co = YinYangSphericalCoords(device, aabb, exp_r=True, N_voxel=scene.n_voxels, r0=scene.r0, interval_th=interval_th)
    reso = co.N_to_reso(scene.n_voxels, aabb)
    assert reso == scene.grid, (reso, scene.grid)
    model = EgoNeRF(aabb, reso, device, co, **scene.model_kwargs())
    model.mlp_mode = mlp_mode
    model.load_state_dict(scene.state_dict, strict=True)
    if scene.emission is not None:
        model.envmap.load_envmap(scene.emission, device)
    model.update_coarse_sigma_grid()
    return model


RENDER_KW = dict(n_coarse=128, n_fine=128, exp_sampling=True, resampling=True, use_coarse_sample=True, interval_th=True,
                 white_bg=False)
