#!/usr/bin/env python
"""CPU emulation of the throughput mode's roundings (which 16-bit format / which stages keep the render inside the
1e-4 rgb bound?).  Uses the oracle (checker code) for the exact path and re-runs its fine pass with the operands of the
fused kernel rounded the way the kernel rounds them:

    app tables -> Q          (render tables stored in Q)
    P = bilinear(plane), L = lerp(line), V = P * L  either in fp32 then rounded to Q, or with every FMA rounded to Q
                             (packed HFMA2 interpolation)
    basis, W1, W2 -> Q; feat / PE input X -> Q; H1 -> Q; accumulation fp32; layer 3 + sigmoid fp32
    density: fp32 tables (exact) or Q tables

Not a test and not on the product path.  Run:  python scripts/error_budget.py [--voxels 128] [--rays 256]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from egonerf_b200.synthetic import make_rays, make_scene  # noqa: E402
from oracle import egn_oracle as O  # noqa: E402


def q(x, dt):
    return x if dt is None else x.to(dt).float()


def taps2d(img, x, y):
    """four taps + weights of F.grid_sample(align_corners=True, zeros) — restated from oracle._tap2d so that the
    accumulation order / rounding can be varied"""
    C, H, W = img.shape
    ix, iy = (x + 1) / 2 * (W - 1), (y + 1) / 2 * (H - 1)
    x0, y0 = torch.floor(ix), torch.floor(iy)
    fx, fy = ix - x0, iy - y0
    x0, y0 = x0.long(), y0.long()
    out = []
    for dx, dy, w in ((0, 0, (1 - fx) * (1 - fy)), (1, 0, fx * (1 - fy)), (0, 1, (1 - fx) * fy), (1, 1, fx * fy)):
        xx, yy = x0 + dx, y0 + dy
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        v = img[:, yy.clamp(0, H - 1), xx.clamp(0, W - 1)].t()
        out.append((v * ok[:, None], w * ok))
    return out


def taps1d(line, t):
    C, L = line.shape
    it = (t + 1) / 2 * (L - 1)
    t0 = torch.floor(it)
    f = it - t0
    t0 = t0.long()
    out = []
    for d, w in ((0, 1 - f), (1, f)):
        tt = t0 + d
        ok = (tt >= 0) & (tt < L)
        out.append((line[:, tt.clamp(0, L - 1)].t() * ok[:, None], w * ok))
    return out


def products(sd, kind, coords, is_yang, table_dt, fma_dt):
    M = coords.shape[0]
    outs = []
    for i in range(3):
        m0, m1 = O.MAT_MODE[i]
        v = O.VEC_MODE[i]
        C = sd[f"{kind}_plane_yin.{i}"].shape[1]
        res = torch.zeros(M, C)
        for h, sel in (("yin", ~is_yang), ("yang", is_yang)):
            if not sel.any():
                continue
            c = coords[sel]
            pt = taps2d(q(sd[f"{kind}_plane_{h}.{i}"][0], table_dt), c[:, m0], c[:, m1])
            lt = taps1d(q(sd[f"{kind}_line_{h}.{i}"][0, :, :, 0], table_dt), c[:, v])
            P = None
            for val, w in pt:
                term = q(w, fma_dt)[:, None] * val
                P = q(term, fma_dt) if P is None else q(P + term, fma_dt)
            Lv = None
            for val, w in lt:
                term = q(w, fma_dt)[:, None] * val
                Lv = q(term, fma_dt) if Lv is None else q(Lv + term, fma_dt)
            res[sel] = q(P * Lv, fma_dt)
        outs.append(res)
    return outs


def render_variant(scene, cfg, rays, aux, dens_dt, app_dt, fma_dt, mlp_dt):
    sd = scene.state_dict
    N = rays.shape[0]
    z, cf, yang = aux["z"], aux["coords"].reshape(-1, 3), aux["is_yang"].reshape(-1)
    S = z.shape[1]
    f = torch.zeros(cf.shape[0])
    for p in products(sd, "density", cf, yang, dens_dt, None):
        f = f + torch.relu(p.sum(-1))
    sigma = O.feature_to_density(f, cfg.density_shift, cfg.fea2dense).view(N, S)
    dist = z[:, 1:] - z[:, :-1]
    dist = torch.cat([dist, dist[:, -1:]], -1)
    alpha, w, bg = O.alpha_composite_weights(sigma, dist * cfg.distance_scale)
    V = q(torch.cat(products(sd, "app", cf, yang, app_dt, fma_dt), -1), mlp_dt)
    feat = torch.zeros(cf.shape[0], 27)
    for h, sel in (("yin", ~yang), ("yang", yang)):
        if sel.any():
            feat[sel] = V[sel] @ q(sd[f"basis_mat_{h}.weight"], mlp_dt).t()
    dirs = rays[:, None, 3:6].expand(N, S, 3).reshape(-1, 3)
    x = torch.cat([feat, dirs, O.freq_encode(feat, cfg.fea_pe), O.freq_encode(dirs, cfg.view_pe)], -1)
    h1 = torch.relu(q(x, mlp_dt) @ q(sd["renderModule.mlp.0.weight"], mlp_dt).t() + q(sd["renderModule.mlp.0.bias"], mlp_dt))
    h2 = torch.relu(q(h1, mlp_dt) @ q(sd["renderModule.mlp.2.weight"], mlp_dt).t() + sd["renderModule.mlp.2.bias"])
    c = torch.sigmoid(h2 @ sd["renderModule.mlp.4.weight"].t() + sd["renderModule.mlp.4.bias"]).view(N, S, 3)
    return torch.sum(w[..., None] * c, -2).clamp(0, 1), alpha


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--voxels", type=int, default=128)
    ap.add_argument("--rays", type=int, default=256)
    ap.add_argument("--smooth", type=int, default=8)
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    scene = make_scene(n_voxels=a.voxels ** 3 if a.voxels != 300 else 27e6, smooth=a.smooth)
    rays = make_rays(a.rays, 'isotropic', seed=77)
    cfg = O.OracleCfg(aabb=scene.aabb, grid=tuple(scene.grid), r0=scene.r0, near=scene.near_far[0], far=scene.near_far[1],
                      density_shift=scene.density_shift, distance_scale=scene.distance_scale)
    with torch.no_grad():
        (rgb, _, _, _, alpha), aux = O.render(scene.state_dict, cfg, rays, False, want_aux=True)
        bf, hf = torch.bfloat16, torch.float16
        variants = [
            ("exact re-run (sanity)", None, None, None, None),
            ("r01 throughput: all tables bf16, fp32 interp, bf16 MMA", bf, bf, None, bf),
            ("density fp32; app bf16 tables, fp32 interp, bf16 MMA", None, bf, None, bf),
            ("density fp32; app bf16 tables, bf16 HFMA2 interp, bf16 MMA", None, bf, bf, bf),
            ("density fp32; app fp32 tables, fp32 interp, bf16 MMA", None, None, None, bf),
            ("density fp32; app fp16 tables, fp32 interp, fp16 MMA", None, hf, None, hf),
            ("density fp32; app fp16 tables, fp16 HFMA2 interp, fp16 MMA", None, hf, hf, hf),
            ("density fp32; app fp32 tables, fp32 interp, fp16 MMA", None, None, None, hf),
            ("density fp16; app fp16 tables, fp16 HFMA2 interp, fp16 MMA", hf, hf, hf, hf),
        ]
        print(f"scene {scene.grid}, {a.rays} rays, smooth={a.smooth}")
        for name, dd, ad, fd, md in variants:
            r, al = render_variant(scene, cfg, rays, aux, dd, ad, fd, md)
            e = (r - rgb).abs()
            print(f"{name:62s} rgb Linf {e.max().item():.2e}  mean {e.mean().item():.2e}  alpha Linf {(al - alpha).abs().max().item():.2e}")


if __name__ == "__main__":
    main()
