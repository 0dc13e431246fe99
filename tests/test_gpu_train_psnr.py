"""GPU: 'PSNR delta vs reference' (BASELINE.json metric) on a short training run.

A student field is fitted to renders of a teacher scene with Adam (learning rates of configs/EgoNeRF/common.txt:
0.02 factors / 0.001 networks), once by autograd through the CPU oracle (= the reference's algorithm, pinned by
tests/golden) and once through libegn_b200 — identical initialisation, identical ray batches, identical injected
jitter uniforms.  After K steps both students are rendered on held-out rays; PSNR = -10 log10(mse) (renderer.py:156-157).
north_star bound: |PSNR_b200 - PSNR_reference| <= 0.05 dB."""
import numpy as np
import pytest
import torch

from tests.helpers import oracle_cfg, scene_for

pytestmark = pytest.mark.gpu
STEPS, BATCH = 30, 256


def _psnr(a, b):
    return float(-10.0 * torch.log10(((a - b) ** 2).mean()))


def _batches(seed):
    from egonerf_b200.synthetic import make_rays
    g = torch.Generator().manual_seed(seed)
    for i in range(STEPS):
        yield make_rays(BATCH, 'isotropic', seed=1000 + i), torch.rand(BATCH, 128, generator=g), torch.rand(BATCH, 128, generator=g)


def _train_gpu(student, teacher_rgb_fn, mode, tables="f32"):
    from egonerf_b200.scene_io import model_from_scene, RENDER_KW
    model = model_from_scene(student, "cuda:0")
    model.mlp_mode, model.table_dtype = mode, tables
    opt = torch.optim.Adam(model.get_optparam_groups(0.02, 0.001), betas=(0.9, 0.99))
    for rays, u_c, u_f in _batches(5):
        target = teacher_rgb_fn(rays).cuda()
        opt.zero_grad(set_to_none=True)
        rgb = model(rays.cuda(), is_train=True, u_coarse=u_c.cuda(), u_fine=u_f.cuda(), **RENDER_KW)[0]
        loss = ((rgb - target) ** 2).mean()
        loss.backward()
        opt.step()
        model.update_coarse_sigma_grid()
    return model


def test_training_psnr_matches_reference_algorithm():
    from egonerf_b200.scene_io import model_from_scene, RENDER_KW
    from egonerf_b200.synthetic import make_rays, make_scene
    from oracle import egn_oracle as O
    teacher = scene_for(dict(n_voxels=40 ** 3, seed=7))
    student = make_scene(n_voxels=40 ** 3, seed=11, sigma_std=0.4)
    cfg = oracle_cfg(teacher)
    tmodel = model_from_scene(teacher, "cuda:0")
    tmodel.mlp_mode = "fp32"

    def teacher_rgb(rays):
        with torch.no_grad():
            return tmodel(rays.cuda(), is_train=False, **RENDER_KW)[0].cpu()

    # reference algorithm: autograd through the CPU oracle
    sd = {k: v.clone().requires_grad_(True) for k, v in student.state_dict.items()}
    fac = [v for k, v in sd.items() if "plane" in k or "line" in k]
    net = [v for k, v in sd.items() if not ("plane" in k or "line" in k)]
    opt = torch.optim.Adam([{"params": fac, "lr": 0.02}, {"params": net, "lr": 0.001}], betas=(0.9, 0.99))
    for rays, u_c, u_f in _batches(5):
        target = teacher_rgb(rays)
        opt.zero_grad(set_to_none=True)
        rgb = O.render(sd, cfg, rays, True, u_c, u_f)[0]
        ((rgb - target) ** 2).mean().backward()
        opt.step()
    held = make_rays(512, 'isotropic', seed=4242)
    gt = teacher_rgb(held)
    with torch.no_grad():
        psnr_ref = _psnr(O.render({k: v.detach() for k, v in sd.items()}, cfg, held, False)[0], gt)
        psnr_init = _psnr(O.render(student.state_dict, cfg, held, False)[0], gt)

    results = {}
    for mode, tables in (("fp32", "f32"), ("tc_split", "f32"), ("tc_bf16", "f32"), ("tc_bf16", "bf16")):
        m = _train_gpu(student, teacher_rgb, mode, tables)
        with torch.no_grad():
            m.mlp_mode, m.table_dtype = "fp32", "f32"           # evaluate every student with the exact renderer
            results[(mode, tables)] = _psnr(m(held.cuda(), is_train=False, **RENDER_KW)[0].cpu(), gt)
    print(f"PSNR on held-out rays after {STEPS} Adam steps (init {psnr_init:.3f} dB): reference algorithm {psnr_ref:.3f} dB; "
          + ", ".join(f"{k[0]}/{k[1]} {v:.3f} dB (delta {v - psnr_ref:+.3f})" for k, v in results.items()))
    assert psnr_ref > psnr_init + 0.5, "the run must actually learn something"
    assert abs(results[("fp32", "f32")] - psnr_ref) <= 0.05
    assert abs(results[("tc_split", "f32")] - psnr_ref) <= 0.05
    assert abs(results[("tc_bf16", "f32")] - psnr_ref) <= 0.05
    assert abs(results[("tc_bf16", "bf16")] - psnr_ref) <= 0.05


def test_long_run_psnr_of_the_default_mode(capsys):
    """128^3 grid, 300 Adam steps of 4096 rays (the run of scripts/train_demo.py, promoted to a test): the default mode
    (fused tcgen05 forward, tcgen05 backward, table-space Adam) against the fp32-equivalent mode (tc_split forward, exact
    fp32 backward, torch Adam) from identical initialisation, batches and sampler seeds.  Every student is evaluated on 8192
    held-out rays with the exact renderer AND with the default-mode renderer.  Training runs are not bit-reproducible (fp32
    atomics in the gradient scatter), so the fp32-equivalent run is made twice and its own run-to-run spread is added to the
    north_star bound of 0.05 dB."""
    from egonerf_b200.optim import TableAdam
    from egonerf_b200.scene_io import model_from_scene, RENDER_KW
    from egonerf_b200.synthetic import make_rays, make_scene
    dev, steps, batch = "cuda:0", 300, 4096
    teacher = model_from_scene(make_scene(n_voxels=128 ** 3, seed=7), dev)
    teacher.mlp_mode = "tc_split"
    student_scene = make_scene(n_voxels=128 ** 3, seed=11, sigma_std=0.4)
    held = make_rays(8192, 'isotropic', seed=4242).to(dev)
    rays_all = make_rays(batch * 64, 'isotropic', seed=99).to(dev)
    with torch.no_grad():
        gt = teacher(held, is_train=False, **RENDER_KW)[0]
        tgt_all = torch.cat([teacher(rays_all[i:i + 65536], is_train=False, **RENDER_KW)[0] for i in range(0, rays_all.shape[0], 65536)])

    def fit(mode, tables, table_adam):
        model = model_from_scene(student_scene, dev)
        model.mlp_mode, model.table_dtype = mode, tables
        opt = TableAdam(model, 0.02, 0.001) if table_adam else \
            torch.optim.Adam(model.get_optparam_groups(0.02, 0.001), betas=(0.9, 0.99), fused=True)
        g = torch.Generator(device=dev).manual_seed(1)
        for it in range(steps):
            idx = torch.randint(0, rays_all.shape[0], (batch,), device=dev, generator=g)
            opt.zero_grad()
            rgb = model(rays_all[idx], is_train=True, seed=1000 + it, **RENDER_KW)[0]
            ((rgb - tgt_all[idx]) ** 2).mean().backward()
            opt.step()
            model.update_coarse_sigma_grid()
        out = {}
        with torch.no_grad():
            for ev_mode, ev_tables in (("tc_split", "f32"), ("tc_f16", "bf16")):
                model.mlp_mode, model.table_dtype = ev_mode, ev_tables
                out[ev_mode] = _psnr(model(held, is_train=False, **RENDER_KW)[0], gt)
        return out

    exact_a = fit("tc_split", "f32", False)
    exact_b = fit("tc_split", "f32", False)
    fast = fit("tc_f16", "bf16", True)
    spread = abs(exact_a["tc_split"] - exact_b["tc_split"])
    ref = 0.5 * (exact_a["tc_split"] + exact_b["tc_split"])
    with capsys.disabled():
        print(f"\n128^3 / 300 steps: fp32-equivalent runs {exact_a['tc_split']:.3f} / {exact_b['tc_split']:.3f} dB (spread {spread:.3f}); "
              f"default mode {fast['tc_split']:.3f} dB with the exact renderer, {fast['tc_f16']:.3f} dB with its own renderer; "
              f"exact student under the default renderer {exact_a['tc_f16']:.3f} dB")
    assert ref > 25.0, "the run must actually learn something (starts at ~10.5 dB)"
    assert abs(fast["tc_split"] - ref) <= 0.05 + spread
    assert abs(fast["tc_f16"] - fast["tc_split"]) <= 0.01          # renderer-to-renderer difference on the same student
    assert abs(exact_a["tc_f16"] - exact_a["tc_split"]) <= 0.01
