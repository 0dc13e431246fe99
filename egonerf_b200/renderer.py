"""Render driver mirror (reference: renderer.py:11-79): chunk loop -> model(...) -> concatenation of the five
outputs.  `OctreeRender_trilinear_fast` is the upstream-TensoRF name BASELINE.json's north_star uses for it."""
import time

import numpy as np
import torch


MIN_EVAL_CHUNK = 65536


def volume_renderer(rays, model, chunk=4096, n_coarse=-1, n_fine=0, ndc_ray=False, white_bg=True, is_train=False,
                    exp_sampling=False, device='cuda', empty_gpu_cache=False, pretrain_envmap=False,
                    pivotal_sample_th=0., resampling=False, use_coarse_sample=True, interval_th=False):
    if pretrain_envmap:
        return model(rays_chunk=rays.to(device), pretrain_envmap=True)
    rgbs, depths, bgs, envs, alphas = [], [], [], [], []
    n_all = rays.shape[0]
    # `chunk` is a memory knob of the reference (its forward materialises (N, S, 150) tensors).  Here a ray's result does
    # not depend on the chunk it is rendered in (tests/test_gpu_parity.py::test_edge_sizes, test_gpu_fullsize.py) and the
    # eval workspace is 24-136 B/sample, so evaluation regroups small chunks (renderer.evaluation passes 4096) into chunks
    # of MIN_EVAL_CHUNK rays: same outputs, 16x fewer launches.  Not for the uniform march, whose eval mode hands on the
    # depths of the FIRST ray of each chunk (EgoNeRF.py:515-516) and is therefore chunk-dependent in the reference itself.
    if not is_train and exp_sampling and chunk < MIN_EVAL_CHUNK:
        chunk = MIN_EVAL_CHUNK
    start = time.time()
    has_env = False
    for c0 in range(0, n_all, chunk):
        rays_chunk = rays[c0:c0 + chunk].to(device, non_blocking=True)
        rgb, depth, bg, env, alpha = model(
            rays_chunk, is_train=is_train, white_bg=white_bg, ndc_ray=ndc_ray, n_coarse=n_coarse, n_fine=n_fine,
            exp_sampling=exp_sampling, pivotal_sample_th=pivotal_sample_th, resampling=resampling,
            use_coarse_sample=use_coarse_sample, interval_th=interval_th, ray_index0=c0)
        has_env = env is not None
        if empty_gpu_cache:          # renderer.py:39-53: every chunk leaves the device as numpy
            rgb, depth, alpha = rgb.cpu().numpy(), depth.cpu().numpy(), alpha.cpu().numpy()
            if has_env:
                bg, env = bg.cpu().numpy(), env.cpu().numpy()
        rgbs.append(rgb); depths.append(depth); alphas.append(alpha)
        if has_env:
            bgs.append(bg); envs.append(env)
    if not is_train:
        print(f"elapsed time per image: {time.time() - start}")
    cat = np.concatenate if empty_gpu_cache else torch.cat
    if has_env:
        return cat(rgbs), cat(depths), cat(bgs), cat(envs), cat(alphas)
    return cat(rgbs), cat(depths), None, None, cat(alphas)


OctreeRender_trilinear_fast = volume_renderer
