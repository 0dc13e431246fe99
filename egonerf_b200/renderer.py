"""Render driver mirror (reference: renderer.py:11-79): chunk loop -> model(...) -> concatenation of the five
outputs.  `OctreeRender_trilinear_fast` is the upstream-TensoRF name BASELINE.json's north_star uses for it."""
import time

import numpy as np
import torch


MIN_EVAL_CHUNK = 65536
HOST_SINK_CHUNK = 16384


class _HostSink:
    """`empty_gpu_cache=True` (renderer.py:39-53: every chunk leaves the device as numpy): instead of one blocking pageable
    `.cpu()` per output and chunk followed by `np.concatenate`, the outputs of the whole call are allocated ONCE in pinned
    host memory (torch's caching host allocator) in their final, concatenated layout, and every chunk is copied into its
    slice asynchronously on a side stream -- the device->host copy of chunk c (1 KB of alpha per ray) overlaps the kernels of
    chunk c+1.  The numpy arrays returned are views of those buffers."""

    def __init__(self, n_all, device):
        self.n_all, self.bufs, self.pending = n_all, None, []
        self.stream = torch.cuda.Stream(device=device)

    def push(self, c0, tensors):
        if self.bufs is None:
            self.bufs = [None if t is None else torch.empty((self.n_all,) + tuple(t.shape[1:]), dtype=t.dtype, pin_memory=True)
                         for t in tensors]
        ev = torch.cuda.Event()
        ev.record()
        self.stream.wait_event(ev)
        with torch.cuda.stream(self.stream):
            for buf, t in zip(self.bufs, tensors):
                if t is not None:
                    buf[c0:c0 + t.shape[0]].copy_(t, non_blocking=True)
                    t.record_stream(self.stream)

    def result(self):
        self.stream.synchronize()
        return [None if b is None else b.numpy() for b in self.bufs]


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
    # When every chunk leaves the device (`empty_gpu_cache=True`, 1 KB of alpha per ray), chunks of HOST_SINK_CHUNK rays let the
    # device->host copy of chunk c run behind the kernels of chunk c+1 instead of after the only chunk of a 65 536-ray call.
    if not is_train and exp_sampling and empty_gpu_cache:
        chunk = min(chunk, HOST_SINK_CHUNK)
    start = time.time()
    has_env = False
    sink = None
    for c0 in range(0, n_all, chunk):
        rays_chunk = rays[c0:c0 + chunk].to(device, non_blocking=True)
        rgb, depth, bg, env, alpha = model(
            rays_chunk, is_train=is_train, white_bg=white_bg, ndc_ray=ndc_ray, n_coarse=n_coarse, n_fine=n_fine,
            exp_sampling=exp_sampling, pivotal_sample_th=pivotal_sample_th, resampling=resampling,
            use_coarse_sample=use_coarse_sample, interval_th=interval_th, ray_index0=c0)
        has_env = env is not None
        if empty_gpu_cache:          # renderer.py:39-53: every chunk leaves the device as numpy
            if sink is None:
                sink = _HostSink(n_all, rgb.device)
            sink.push(c0, [rgb.detach(), depth.detach(), bg.detach() if has_env else None, env.detach() if has_env else None,
                           alpha.detach()])
            continue
        rgbs.append(rgb); depths.append(depth); alphas.append(alpha)
        if has_env:
            bgs.append(bg); envs.append(env)
    if empty_gpu_cache and sink is not None:
        out = sink.result()
        if not is_train:
            print(f"elapsed time per image: {time.time() - start}")
        return out[0], out[1], out[2], out[3], out[4]
    if not is_train:
        print(f"elapsed time per image: {time.time() - start}")
    if has_env:
        return torch.cat(rgbs), torch.cat(depths), torch.cat(bgs), torch.cat(envs), torch.cat(alphas)
    return torch.cat(rgbs), torch.cat(depths), None, None, torch.cat(alphas)


OctreeRender_trilinear_fast = volume_renderer
