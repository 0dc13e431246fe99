#!/usr/bin/env python
"""Sample locality of the fine pass (VERDICT r01 item 5 / north_star "TMA-staged factor tiles"): for the rays of a bench
workload, where do the 18 taps of consecutive samples land?

For every 128-sample tile (half a ray at S = 256; the unit `egn_fused_fine_kernel` processes) it reports
  * how many samples share the tile's most frequent (hemisphere, theta-cell, phi-cell) pair -> could read planes 0/1 from
    two staged r-rows and plane 2 / lines 0,1 from registers,
  * the r-cell span of those samples (the window a staged strip would have to cover),
  * per tap family, the fraction of samples whose clamped texel indices equal the PREVIOUS sample's (tap reuse in registers).
Writes gpurun_out/strip_histogram_<workload>.json and prints a summary.  Needs a GPU (sampler + coordinates run in libegn_b200).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import WORKLOADS  # noqa: E402
from egonerf_b200.scene_io import model_from_scene  # noqa: E402
from egonerf_b200.synthetic import make_rays, make_scene  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--rays", type=int, default=8192)
    ap.add_argument("--tile", type=int, default=128)
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    dev = torch.device("cuda", 0)
    scene = make_scene(n_voxels=wl["n_voxels"], **wl["scene"])
    model = model_from_scene(scene, dev)
    nc, nf = wl.get("n_coarse", 128), wl.get("n_fine", 128)
    if wl.get("kind") == "erp":
        rays = make_rays(2048 * 4096, 'erp', erp_hw=(2048, 4096))
        rays = rays[torch.randperm(rays.shape[0], generator=torch.Generator().manual_seed(3))[:args.rays]]
    else:
        rays = make_rays(args.rays, 'isotropic', seed=1000)
    rays = rays.to(dev)
    z = model.sample_depths(rays, is_train=False, n_coarse=nc, n_fine=nf)                   # (N, S) sorted
    N, S = z.shape
    pts = rays[:, None, :3] + rays[:, None, 3:] * z[..., None]
    c7 = model.coordinates.cart_to_normalized(pts)                                         # (N, S, 7)
    yang = c7[..., 6] != 0
    c = torch.where(yang[..., None], c7[..., 3:6], c7[..., 0:3])
    G = torch.tensor(model.gridSize.tolist(), device=dev)
    ix = (c + 1) / 2 * (G - 1).float()
    cell = torch.floor(ix).long().clamp(min=-1)
    cell = torch.minimum(cell, G)                                                          # (N, S, 3): r, theta, phi cells
    T = args.tile
    nt = S // T
    cell_t = cell.view(N, nt, T, 3)
    yang_t = yang.view(N, nt, T)
    # ---- mode of (hemisphere, theta-cell, phi-cell) per tile ----
    key = (yang_t.long() * 4096 + cell_t[..., 1] + 1) * 4096 + cell_t[..., 2] + 1
    mode = torch.mode(key, dim=-1).values
    share = (key == mode[..., None])
    n_share = share.sum(-1)                                                                # (N, nt)
    big = torch.full_like(cell_t[..., 0], 1 << 30)
    rmin = torch.where(share, cell_t[..., 0], big).amin(-1)
    rmax = torch.where(share, cell_t[..., 0], -big).amax(-1)
    span = (rmax - rmin + 2).clamp(min=0)                                                  # texels of one staged r-row
    # same hemisphere + same theta-cell only (plane 0 rows) / same phi-cell only (plane 1 rows)
    def mode_share(k):
        m = torch.mode(k, dim=-1).values
        return (k == m[..., None]).sum(-1)
    th_share = mode_share(yang_t.long() * 4096 + cell_t[..., 1] + 1)
    ph_share = mode_share(yang_t.long() * 4096 + cell_t[..., 2] + 1)
    # ---- reuse against the previous sample of the same ray ----
    same = lambda *ax: ((cell[:, 1:, list(ax)] == cell[:, :-1, list(ax)]).all(-1) & (yang[:, 1:] == yang[:, :-1])).float().mean().item()
    res = {
        "workload": args.workload, "rays": N, "samples_per_ray": S, "tile": T, "grid": G.tolist(),
        "tile_share_theta_phi": {"mean": n_share.float().mean().item() / T,
                                 "tiles_ge_50pct": (n_share >= T // 2).float().mean().item(),
                                 "tiles_ge_90pct": (n_share >= int(0.9 * T)).float().mean().item(),
                                 "by_tile_index": [n_share[:, i].float().mean().item() / T for i in range(nt)]},
        "tile_share_theta_only": th_share.float().mean().item() / T,
        "tile_share_phi_only": ph_share.float().mean().item() / T,
        "r_span_texels_of_sharing_samples": {"mean": span.float().mean().item(), "p50": span.float().median().item(),
                                             "p90": span.float().quantile(0.9).item(), "max": span.max().item()},
        "distinct_r_cells_per_tile": torch.tensor([[len(torch.unique(cell_t[i, t, :, 0])) for t in range(nt)]
                                                   for i in range(min(N, 512))]).float().mean().item(),
        "prev_sample_reuse": {"plane0 (r,theta)": same(0, 1), "plane1 (r,phi)": same(0, 2), "plane2 (theta,phi)": same(1, 2),
                              "line0 (phi)": same(2), "line1 (theta)": same(1), "line2 (r)": same(0),
                              "all 18 taps": same(0, 1, 2)},
        "yang_fraction": yang.float().mean().item(),
        "hemisphere_switches_per_ray": (yang[:, 1:] != yang[:, :-1]).float().sum(-1).mean().item(),
    }
    pr = res["prev_sample_reuse"]
    res["taps_reloaded_per_sample_with_reuse"] = (4 * (1 - pr["plane0 (r,theta)"]) + 4 * (1 - pr["plane1 (r,phi)"]) +
                                                  4 * (1 - pr["plane2 (theta,phi)"]) + 2 * (1 - pr["line0 (phi)"]) +
                                                  2 * (1 - pr["line1 (theta)"]) + 2 * (1 - pr["line2 (r)"]))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"strip_histogram_{args.workload}.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
