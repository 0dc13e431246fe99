"""GPU: the loop body of the reference's train.py (train.py:245-357) and its evaluation call (renderer.py:129-134), restated
against a synthetic dataset and driven THROUGH `shim/` -- the module paths an unmodified train.py imports -- for a few
iterations.  The GPU box has no reference checkout, so the shim runs in its stand-alone mode here; the CPU test
tests/test_shim_cpu.py::test_unmodified_train_py_runs_through_the_shim runs the reference's own `train()` against the same shim."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

LOOP = r'''
import sys, os, math
sys.path.insert(0, sys.argv[1])                       # <repo>/shim
import numpy as np, torch
from renderer import volume_renderer                  # train.py:6
from models.EgoNeRF import EgoNeRF                     # train.py:12
from models import coordinates_dict                    # train.py:15
sys.path.insert(0, os.path.dirname(sys.argv[1]))
from egonerf_b200.synthetic import make_rays
device = torch.device("cuda")
torch.manual_seed(20221028); np.random.seed(20221028)

class TVLoss(torch.nn.Module):                         # utils.py:155-171 restated (the box has no reference tree)
    def forward(self, x):
        b, _, h, w = x.shape
        ch, cw = x[:, :, 1:, :].numel(), x[:, :, :, 1:].numel()
        return 2 * (torch.pow(x[:, :, 1:, :] - x[:, :, :h - 1, :], 2).sum() / ch + torch.pow(x[:, :, :, 1:] - x[:, :, :, :w - 1], 2).sum() / cw) / b

near_far = [0.1, 300.]; half = 0.5 + near_far[1]
aabb = torch.tensor([[-half] * 3, [half] * 3]).to(device)
coordinates = coordinates_dict['yinyang'](device, aabb, exp_r=True, N_voxel=40 ** 3, r0=0.05, interval_th=True)   # train.py:122-124
reso_cur = coordinates.N_to_reso(40 ** 3, aabb)
model = EgoNeRF(aabb, reso_cur, device, coordinates, density_n_comp=[16] * 3, appearance_n_comp=[48] * 3, app_dim=27,
                near_far=near_far, shadingMode='MLP_Fea', alphaMask_thres=1e-4, density_shift=-10, distance_scale=25, pos_pe=6,
                view_pe=2, fea_pe=2, featureC=128, step_ratio=0.5, fea2denseAct='softplus', use_envmap=True, envmap_res_H=64,
                coarse_sigma_grid_update_rule='conv', coarse_sigma_grid_reso=None, interval_th=True)               # train.py:163-171
grad_vars = model.get_optparam_groups(0.02, 0.001, 0.1)
optimizer = torch.optim.Adam(grad_vars, betas=(0.9, 0.99))                                                        # train.py:186
n_iters = 6; lr_factor = 0.1 ** (1 / n_iters)
allrays = make_rays(4096, 'isotropic', seed=3); allrgbs = torch.rand(4096, 3, generator=torch.Generator().manual_seed(4))
tvreg = TVLoss(); TV_weight_density, TV_weight_app = 0.1, 0.01                                                    # ricoh/common.txt:12-13
losses = []
for iteration in range(n_iters):
    ray_idx = torch.randint(0, 4096, (1024,))
    rays_train, rgb_train = allrays[ray_idx], allrgbs[ray_idx].to(device)
    rgb_map, depth_map, _, _, alpha = volume_renderer(rays_train, model, chunk=1024, n_coarse=128, n_fine=128, white_bg=False,
        ndc_ray=False, device=device, is_train=True, exp_sampling=True, pivotal_sample_th=0., resampling=True,
        use_coarse_sample=True, interval_th=True)                                                                # train.py:253-258
    loss = torch.mean((rgb_map - rgb_train) ** 2)
    total_loss = loss
    TV_weight_density *= lr_factor
    total_loss = total_loss + model.TV_loss_density(tvreg) * TV_weight_density                                   # train.py:293-297
    TV_weight_app *= lr_factor
    total_loss = total_loss + model.TV_loss_app(tvreg) * TV_weight_app
    optimizer.zero_grad(); total_loss.backward(); optimizer.step()                                               # train.py:311-313
    losses.append(loss.detach().item())
    for param_group in optimizer.param_groups:
        param_group['lr'] = param_group['lr'] * lr_factor
    model.update_coarse_sigma_grid()                                                                              # train.py:356-357
assert all(math.isfinite(l) for l in losses), losses
assert losses[-1] < losses[0], losses
assert tuple(alpha.shape) == (1024, 257) and depth_map.shape == (1024,)
g = model.density_plane_yin[0].grad
assert g is not None and torch.isfinite(g).all() and float(g.abs().sum()) > 0
path = os.path.join(sys.argv[2], 'dropin.th'); model.save(path, global_step=n_iters)                              # train.py:379
ckpt = torch.load(path, map_location=device, weights_only=False)
kwargs = ckpt['kwargs']; kwargs.update({'device': device})
model2 = EgoNeRF(**kwargs); assert model2.load(ckpt) == n_iters                                                   # train.py:155-160
with torch.no_grad():
    outs = [volume_renderer(allrays[:512], m, chunk=4096, n_coarse=128, n_fine=128, ndc_ray=False, white_bg=False, exp_sampling=True,
                            device=device, empty_gpu_cache=True, resampling=True, use_coarse_sample=True, interval_th=True) for m in (model, model2)]
assert isinstance(outs[0][0], np.ndarray) and outs[0][4].shape == (512, 257)                                      # renderer.py:39-53
assert np.array_equal(outs[0][0], outs[1][0])
print('DROPIN_LOOP_OK', ' '.join(f'{l:.5f}' for l in losses))
'''


@pytest.mark.gpu
def test_train_loop_body_through_the_shim(tmp_path):
    env = {k: v for k, v in os.environ.items() if k != "EGONERF_REFERENCE"}
    out = subprocess.run([sys.executable, "-c", LOOP, os.path.join(ROOT, "shim"), str(tmp_path)], capture_output=True, text=True,
                         timeout=600, cwd=str(tmp_path), env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    assert "DROPIN_LOOP_OK" in out.stdout
