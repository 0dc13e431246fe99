"""Freeze outputs of the UNMODIFIED reference into tests/golden/*.npz  (run in the build container only).

    python oracle/make_golden.py

TEST INFRASTRUCTURE.  The reference (a Python program) cannot travel to the GPU box, so its results on
seeded synthetic inputs are committed as small fixtures together with this generating script.  Every
fixture records the inputs needed to replay it (scene seed or the tensors themselves, rays, the uniform
draws of train mode) and the reference's outputs.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_harness import import_reference          # noqa: E402
from egonerf_b200.synthetic import make_scene, make_rays  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
coordinates_dict, RefEgoNeRF, ref_volume_renderer, ref_sample_pdf, ref_raw2alpha = import_reference()


def build_reference(scene, interval_th=True):
    co = coordinates_dict['yinyang']('cpu', scene.aabb, exp_r=True, N_voxel=scene.n_voxels, r0=scene.r0,
                                     interval_th=interval_th)
    reso = co.N_to_reso(scene.n_voxels, scene.aabb)
    assert reso == scene.grid, (reso, scene.grid)
    model = RefEgoNeRF(scene.aabb, reso, 'cpu', co, **dict(scene.model_kwargs(), interval_th=interval_th))
    model.load_state_dict(scene.state_dict, strict=True)
    if scene.emission is not None:
        model.envmap.emission = scene.emission.clone().requires_grad_(True)
    model.update_coarse_sigma_grid()
    return co, model


def ref_render(model, rays, is_train, n_coarse=128, n_fine=128, resampling=True, use_coarse_sample=True, exp_sampling=True,
               interval_th=True):
    return ref_volume_renderer(rays, model, chunk=rays.shape[0], n_coarse=n_coarse, n_fine=n_fine,
                               is_train=is_train, exp_sampling=exp_sampling, resampling=resampling,
                               use_coarse_sample=use_coarse_sample, interval_th=interval_th, device='cpu',
                               white_bg=False)


def npz(name, **kw):
    arrs = {}
    for k, v in kw.items():
        if v is None:
            continue
        arrs[k] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **arrs)
    print(f"wrote {name}: {os.path.getsize(path) / 1024:.1f} KiB")


def sd_checksum(sd):
    return np.array([float(v.double().sum()) for v in sd.values()] +
                    [float(v.double().abs().sum()) for v in sd.values()])


def main():
    os.makedirs(OUT, exist_ok=True)

    # ---- 1. schedules and coordinate KATs (EgoNeRF.py:68-76, coordinates.py:110-156,442-498) -------
    for tag, nvox, nf, r0 in (("indoor300", 27e6, (0.01, 15.), 0.03), ("indoor128", 128 ** 3, (0.01, 15.), 0.03),
                              ("outdoor300", 27e6, (0.1, 300.), 0.05)):
        sc = make_scene(n_voxels=32 ** 3, near_far=nf, r0=r0)     # aabb only depends on far
        co = coordinates_dict['yinyang']('cpu', sc.aabb, exp_r=True, N_voxel=nvox, r0=r0, interval_th=True)

        class _M:   # minimal object exposing what sample_ray_exp reads
            near_far = list(nf)
            coordinates = co
            aabb = sc.aabb
        rays = make_rays(4, 'probe', seed=3)
        _, z, _ = RefEgoNeRF.sample_ray_exp(_M, rays[:, :3], rays[:, 3:], is_train=False, N_samples=128)
        g = torch.Generator().manual_seed(5)
        pts = torch.cat([
            torch.tensor([[1., 0, 0], [0, 1., 0], [0, 0, 1.], [-1., 0, 0], [.3, -.2, .1], [3., -2, 1], [-5., .5, 7],
                          [10., 10, 10], [0., 0, 0], [0, -1., 0], [0, 0, -1.], [40., 0, 0]]),
            (torch.rand(2000, 3, generator=g) - .5) * 2 * nf[1],
            torch.randn(2000, 3, generator=g) * 0.3])
        unn = co.from_cartesian(pts)
        nrm = co.normalize_coord(unn, downsample=2)
        # the r ladder the reference rebuilds on every call: recover it through normalize_r's inverse property
        npz(f"kat_coords_{tag}.npz", aabb=sc.aabb, grid=np.array(co.N_to_reso(nvox, sc.aabb)), r0=r0, near_far=np.array(nf),
            far_r=co.far[0], z_coarse=z[0], points=pts, unnormalized=unn, normalized=nrm)

    # ---- 2. stand-alone operators on a small grid (params stored) --------------------------------
    small = make_scene(n_voxels=40 ** 3, seed=7)
    co, model = build_reference(small)
    g = torch.Generator().manual_seed(11)
    M = 3000
    c3 = torch.rand(M, 3, generator=g) * 2.2 - 1.1            # includes out-of-range taps
    is_yang = torch.rand(M, generator=g) < 0.5
    coords7 = torch.zeros(M, 7)
    coords7[~is_yang, 0:3] = c3[~is_yang]
    coords7[is_yang, 3:6] = c3[is_yang]
    coords7[:, 6] = is_yang.float()
    with torch.no_grad():
        sig = model.compute_densityfeature(coords7)
        sigc = model.compute_coarse_densityfeature(coords7)
        app = model.compute_appfeature(coords7)
        dirs = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1)
        rgb = model.renderModule(None, dirs, app)
    npz("ops_small.npz", seed=7, n_voxels=40 ** 3, coords7=coords7, sigma_feature=sig, coarse_sigma_feature=sigc,
        app_feature=app, dirs=dirs, rgb=rgb, checksum=sd_checksum(small.state_dict))

    # raw2alpha / sample_pdf (tensorBase.py:22-27, ray_utils.py:156-187)
    sg = torch.rand(16, 128, generator=g) * 0.5 * (torch.rand(16, 128, generator=g) < 0.3)
    ds = torch.rand(16, 128, generator=g) * 2
    a, w, bgw = ref_raw2alpha(sg, ds)
    bins = torch.sort(torch.rand(16, 127, generator=g) * 15, -1)[0]
    fz_eval = ref_sample_pdf(bins, w[:, 1:-1], 128, is_train=False)
    torch.manual_seed(99)
    fz_train = ref_sample_pdf(bins, w[:, 1:-1], 128, is_train=True)
    torch.manual_seed(99)
    u_train = torch.rand(16, 128)
    npz("composite_pdf.npz", sigma=sg, dist=ds, alpha=a, weight=w, bg=bgw, bins=bins, fine_eval=fz_eval,
        fine_train=fz_train, u_train=u_train)

    # ---- 3. whole-path renders ---------------------------------------------------------------------
    def render_case(name, scene, rays, is_train, seed=1234, grads=False, mse=False, **kw):
        co, model = build_reference(scene, kw.get("interval_th", True))
        N = rays.shape[0]
        u_c = u_f = None
        if is_train:
            torch.manual_seed(seed)
            u_c = torch.rand(N, 128)            # EgoNeRF.py:81 rand_like(r)
            u_f = torch.rand(N, 128)            # ray_utils.py:169
            torch.manual_seed(seed)
        store = dict(rays=rays, is_train=int(is_train), u_coarse=u_c, u_fine=u_f)
        # the reference does not return its sorted sample depths (EgoNeRF.py:537): capture them from torch.sort
        captured = []
        real_sort = torch.sort

        def spy_sort(*a, **k):
            r = real_sort(*a, **k)
            captured.append(r[0].detach().clone())
            return r
        torch.sort = spy_sort
        if grads:
            out = ref_render(model, rays, is_train, **kw)
            g2 = torch.Generator().manual_seed(seed + 1)
        if grads and mse:
            # the loss train.py:260 trains with: mean squared error against a target image (no sign cancellation between rays)
            target = torch.rand(out[0].shape, generator=g2)
            loss = torch.mean((out[0] - target) ** 2)
            loss.backward()
            store.update(target=target, loss=loss.detach())
            for k, p in model.named_parameters():
                store["grad:" + k] = p.grad if p.grad is not None else torch.zeros_like(p)
            if scene.emission is not None:
                store["grad:envmap.emission"] = model.envmap.emission.grad
        elif grads:
            wr = torch.randn(out[0].shape, generator=g2)
            wa = torch.randn(out[4].shape, generator=g2) * 0.01
            loss = (out[0] * wr).sum() + (out[4] * wa).sum()
            if out[2] is not None:
                wb = torch.randn(out[2].shape, generator=g2)
                we = torch.randn(out[3].shape, generator=g2)
                loss = loss + (out[2] * wb).sum() + (out[3] * we).sum()
                store.update(w_bg=wb, w_env=we)
            loss.backward()
            store.update(w_rgb=wr, w_alpha=wa, loss=loss.detach())
            for k, p in model.named_parameters():
                store["grad:" + k] = p.grad if p.grad is not None else torch.zeros_like(p)
            if scene.emission is not None:
                store["grad:envmap.emission"] = model.envmap.emission.grad
        else:
            with torch.no_grad():
                out = ref_render(model, rays, is_train, **kw)
        torch.sort = real_sort
        if captured:
            store["z_vals"] = captured[0]
        store.update(rgb=out[0], depth=out[1], bg=out[2], env=out[3], alpha=out[4],
                     checksum=sd_checksum(scene.state_dict))
        npz(name, **store)

    tiny = make_scene(n_voxels=40 ** 3, seed=7)
    tiny_env = make_scene(n_voxels=40 ** 3, seed=8, envmap_h=32, near_far=(0.1, 300.), r0=0.05, density_shift=-10.)
    r64 = make_rays(64, 'isotropic', seed=21)
    r64p = make_rays(64, 'probe', seed=22)
    render_case("render_tiny_eval.npz", tiny, r64, False)
    render_case("render_tiny_train.npz", tiny, r64p, True)
    render_case("render_tiny_env_eval.npz", tiny_env, r64, False)
    render_case("render_tiny_env_train_grad.npz", tiny_env, r64p, True, grads=True)
    render_case("render_tiny_train_grad.npz", tiny, r64, True, grads=True, seed=77)
    render_case("render_tiny_noresample.npz", tiny, r64, False, resampling=False, n_fine=0)
    render_case("render_tiny_fineonly.npz", tiny, r64, False, use_coarse_sample=False)
    # BASELINE.json configs[1] / configs[2] shapes: scene regenerated from its seed, checksum pinned
    render_case("render_128_eval.npz", make_scene(n_voxels=128 ** 3), make_rays(256, 'isotropic', seed=31), False)
    render_case("render_128_white_eval.npz", make_scene(n_voxels=128 ** 3, smooth=1), make_rays(256, 'isotropic', seed=34), False)
    render_case("render_300_eval.npz", make_scene(n_voxels=27e6), make_rays(128, 'isotropic', seed=32), False)
    render_case("render_300_train.npz", make_scene(n_voxels=27e6), make_rays(128, 'isotropic', seed=33), True)
    # uniform march (exp_sampling=False, TensorBase.sample_ray tensorBase.py:308-327): no shipped config uses it; SURVEY a4
    render_case("render_tiny_march_eval.npz", tiny, r64, False, exp_sampling=False)
    render_case("render_tiny_march_train.npz", tiny, r64p, True, exp_sampling=False)
    # other decoders that work through EgoNeRF.forward in the reference: MLP, RGB
    render_case("render_tiny_mlp.npz", make_scene(n_voxels=40 ** 3, seed=9, shading='MLP'), r64, False)
    render_case("render_tiny_rgb.npz", make_scene(n_voxels=40 ** 3, seed=10, shading='RGB', app_dim=3), r64, False)

    # ---- 3b. a run WITHOUT --interval_th (opt.py:190 default): plain exponential ladders (EgoNeRF.py:59-67,
    # coordinates.py:132-156), coarse pass on the N_r/2 ladder
    render_case("render_tiny_plain_eval.npz", tiny, r64, False, interval_th=False)
    render_case("render_tiny_plain_train_grad.npz", tiny, r64p, True, grads=True, seed=78, interval_th=False)
    render_case("render_tiny_train_mse_grad.npz", tiny, r64, True, grads=True, mse=True, seed=91)
    render_case("render_tiny_env_train_mse_grad.npz", tiny_env, r64, True, grads=True, mse=True, seed=92)
    render_case("render_tiny_plain_noresample.npz", tiny, r64, False, resampling=False, n_fine=0, interval_th=False)
    co = coordinates_dict['yinyang']('cpu', tiny.aabb, exp_r=True, N_voxel=tiny.n_voxels, r0=tiny.r0, interval_th=False)
    gk = torch.Generator().manual_seed(6)
    pts = torch.cat([torch.tensor([[1., 0, 0], [0, 0, 1.], [.03, 0, 0], [.02, .01, 0], [0., 0, 0], [26., 0, 0], [30., 0, 0]]),
                     (torch.rand(1500, 3, generator=gk) - .5) * 30, torch.randn(1500, 3, generator=gk) * 0.3])
    unn = co.from_cartesian(pts)
    npz("kat_coords_plain.npz", aabb=tiny.aabb, grid=np.array(co.N_to_reso(tiny.n_voxels, tiny.aabb)), r0=tiny.r0, far_r=co.far[0],
        points=pts, normalized=co.normalize_coord(unn), normalized_coarse=co.normalize_coord(unn, downsample=2))

    # ---- 4. coarse-to-fine upsampling (train.py:371-377): 20^3 -> 28^3 voxels, then a render on the new grid
    up = make_scene(n_voxels=20 ** 3, seed=3)
    co, model = build_reference(up)
    reso = co.N_to_reso(28 ** 3, model.aabb)
    model.upsample_volume_grid(reso)
    co.set_resolution(reso)                       # also resets r0 to 0.05 (coordinates.py:214)
    model.update_coarse_sigma_grid()
    with torch.no_grad():
        out = ref_render(model, r64, False)
    factors = {"sd:" + k: v for k, v in model.state_dict().items() if "plane" in k or "line" in k}
    npz("upsample_tiny.npz", seed=3, n_voxels=20 ** 3, grid_old=np.array(up.grid), grid_new=np.array(reso), r0_after=co.r0,
        rays=r64, rgb=out[0], depth=out[1], alpha=out[4], checksum=sd_checksum(up.state_dict), **factors)

    # ---- 5. occupancy mask (EgoNeRF.py:11-24,438-489, tensorBase.py:421-436; deprecated in the reference) -------
    import warnings
    warnings.simplefilter("ignore", DeprecationWarning)
    co, model = build_reference(tiny)
    gs = (8, 9, 20)
    with torch.no_grad():
        a_yin, a_yang = model.getDenseAlpha(gs)
        dense = torch.cat([a_yin.reshape(-1), a_yang.reshape(-1)])
        pooled = torch.cat([torch.nn.functional.max_pool3d(a.clamp(0, 1).transpose(0, 2).contiguous()[None, None], 3, 1, 1).reshape(-1)
                            for a in (a_yin, a_yang)])
        # a threshold that keeps about half of the lattice and that no pooled alpha sits within 1e-4 of
        srt = torch.sort(pooled.unique())[0]
        gaps = srt[1:] - srt[:-1]
        k = int(torch.argmin((srt[:-1] - pooled.median()).abs() + (gaps < 1e-3) * 1e3))
        thres = float((srt[k] + srt[k + 1]) / 2)
        assert (pooled - thres).abs().min() > 1e-4
        model.alphaMask_thres = thres
        model.updateAlphaMask(gs)
        g5 = torch.Generator().manual_seed(12)
        M = 2000
        c3 = torch.rand(M, 3, generator=g5) * 2.1 - 1.05
        yang = torch.rand(M, generator=g5) < 0.5
        c7 = torch.zeros(M, 7)
        c7[~yang, 0:3] = c3[~yang]
        c7[yang, 3:6] = c3[yang]
        c7[:, 6] = yang.float()
        masked_alpha = model.compute_alpha(c7, model.stepSize)
    npz("alpha_mask_tiny.npz", grid=np.array(gs), thres=thres, step=model.stepSize, alpha_yin=a_yin, alpha_yang=a_yang,
        mask_yin=model.alphaMask.alpha_volume_yin, mask_yang=model.alphaMask.alpha_volume_yang, coords7=c7,
        masked_alpha=masked_alpha, checksum=sd_checksum(tiny.state_dict))
    print("mask occupancy", float(model.alphaMask.alpha_volume_yin.mean()), float(model.alphaMask.alpha_volume_yang.mean()),
          "rejected samples", float((masked_alpha == 0).float().mean()))

    # ---- 6. regularisers (utils.py:155-183, EgoNeRF.py:189-229): values and the gradient of the weighted sum the
    # Ricoh configs train with (train.py:288-305; ricoh/common.txt:12-13 TV 0.1 / 0.01) plus an L1 term, w.r.t. all 24 factors
    from utils import TVLoss, ray_entropy_loss                       # the reference's own (import_reference left it in sys.modules)
    co, model = build_reference(tiny)
    tv = TVLoss()
    vals = [model.TV_loss_density(tv), model.TV_loss_app(tv), model.density_L1(), model.vector_comp_diffs()]
    w_reg = (0.1, 0.01, 0.05)
    (w_reg[0] * vals[0] + w_reg[1] * vals[1] + w_reg[2] * vals[2]).backward()
    g6 = torch.Generator().manual_seed(21)
    alpha_in = torch.rand(64, 257, generator=g6)
    reg_grads = {"grad:" + k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in model.named_parameters()
                 if "plane" in k or "line" in k}
    npz("regularisers_tiny.npz", weights=np.array(w_reg), values=torch.stack([v.detach() for v in vals]), alpha=alpha_in,
        ray_entropy=ray_entropy_loss(alpha_in), checksum=sd_checksum(tiny.state_dict), **reg_grads)

    # ---- 7. equirectangular rays (dataLoader/ray_utils.py:24-40,85-113 + dataset_omniblender.py:42-43) -------
    from dataLoader.ray_utils import get_ray_directions_360, get_rays
    H, W = 24, 48
    g7 = torch.Generator().manual_seed(31)
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g7))
    c2w = torch.cat([q, torch.tensor([[0.3], [-0.1], [0.2]])], 1).float()
    directions = get_ray_directions_360(H, W)
    directions = directions / torch.norm(directions, dim=-1, keepdim=True)
    rays_o, rays_d = get_rays(directions, c2w)
    npz("erp_rays_ref.npz", H=H, W=W, c2w=c2w, rays=torch.cat([rays_o, rays_d], 1))


if __name__ == "__main__":
    main()
