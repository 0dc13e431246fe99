"""Teacher-student fit on the GPU box: a student field (default-style init) is trained on renders of a synthetic teacher
scene with the reference's optimiser settings (Adam, lr 0.02 / 0.001, betas (0.9, 0.99), train.py:172-186), once in the
fp32-parity mode and once in the throughput mode (fused tcgen05 kernels, bf16 tables), from identical initialisation,
ray batches and sampler seeds.  Prints PSNR on held-out rays (renderer.py:156-157) every few steps; every student is
evaluated with the exact renderer.

    python scripts/train_demo.py [--voxels 2097152] [--steps 300] [--batch 4096]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from egonerf_b200.scene_io import RENDER_KW, model_from_scene          # noqa: E402
from egonerf_b200.synthetic import make_rays, make_scene                # noqa: E402


def psnr(a, b):
    return float(-10.0 * torch.log10(((a - b) ** 2).mean()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--voxels", type=float, default=128 ** 3)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--every", type=int, default=50)
    ap.add_argument("--table-adam", action="store_true", help="throughput run uses egonerf_b200.optim.TableAdam")
    args = ap.parse_args()
    dev = "cuda:0"
    teacher = model_from_scene(make_scene(n_voxels=args.voxels, seed=7), dev)
    teacher.mlp_mode = "tc_split"
    student_scene = make_scene(n_voxels=args.voxels, seed=11, sigma_std=0.4)
    held = make_rays(8192, 'isotropic', seed=4242).to(dev)
    with torch.no_grad():
        gt = teacher(held, is_train=False, **RENDER_KW)[0]
    rays_all = make_rays(args.batch * 64, 'isotropic', seed=99).to(dev)
    with torch.no_grad():
        tgt_all = torch.cat([teacher(rays_all[i:i + 65536], is_train=False, **RENDER_KW)[0] for i in range(0, rays_all.shape[0], 65536)])
    out = {}
    for name, mode, tables in (("parity (tc_split fwd, fp32 bwd)", "tc_split", "f32"), ("throughput (tc_bf16, bf16 tables)", "tc_bf16", "bf16")):
        model = model_from_scene(student_scene, dev)
        model.mlp_mode, model.table_dtype = mode, tables
        if args.table_adam and mode == "tc_bf16":
            from egonerf_b200.optim import TableAdam
            opt = TableAdam(model, 0.02, 0.001)
        else:
            opt = torch.optim.Adam(model.get_optparam_groups(0.02, 0.001), betas=(0.9, 0.99), fused=True)
        g = torch.Generator(device=dev).manual_seed(1)
        curve = []
        torch.cuda.synchronize()
        t0 = time.time()
        for it in range(args.steps + 1):
            if it % args.every == 0:
                with torch.no_grad():
                    m0, t0m = model.mlp_mode, model.table_dtype
                    model.mlp_mode, model.table_dtype = "tc_split", "f32"
                    curve.append((it, round(psnr(model(held, is_train=False, **RENDER_KW)[0], gt), 3)))
                    model.mlp_mode, model.table_dtype = m0, t0m
            if it == args.steps:
                break
            idx = torch.randint(0, rays_all.shape[0], (args.batch,), device=dev, generator=g)
            opt.zero_grad()
            rgb = model(rays_all[idx], is_train=True, seed=1000 + it, **RENDER_KW)[0]
            loss = ((rgb - tgt_all[idx]) ** 2).mean()
            loss.backward()
            opt.step()
            model.update_coarse_sigma_grid()
        torch.cuda.synchronize()
        out[name] = {"psnr_curve": curve, "seconds": round(time.time() - t0, 2)}
        print(name, curve, f"{time.time() - t0:.1f} s")
    a, b = list(out.values())
    print("final PSNR delta (throughput - parity): %+.3f dB" % (b["psnr_curve"][-1][1] - a["psnr_curve"][-1][1]))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
