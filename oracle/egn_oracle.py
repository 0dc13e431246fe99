"""CPU oracle: a plain restatement of the EgoNeRF volume-rendering path (reference = changwoonchoi/EgoNeRF).

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this file, and only as the checker /
the CPU baseline.  The product path (`egonerf_b200/`) never imports it and has no CPU fallback.

Parity pin: the reference ships no tests and no golden vectors ("parity unpinned" by the reference
itself).  The pin used instead: `oracle/make_golden.py` runs the UNMODIFIED reference
(imported from /root/reference in the build container, see `oracle/ref_harness.py`) and freezes its
inputs/outputs in `tests/golden/*.npz`; `tests/test_oracle_golden.py` holds this restatement to those
vectors.  Every function cites the reference lines it follows.

Arithmetic is fp32 torch-CPU, written sample-by-sample with explicit bilinear taps (no F.grid_sample,
no boolean-mask compaction) so that it is an independent statement of the semantics:
  state dict keys and (1,C,H,W) shapes are the reference's (`EgoNeRF.init_one_svd`, EgoNeRF.py:102-122).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch

PI = math.pi


# --------------------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------------------
@dataclass
class OracleCfg:
    aabb: torch.Tensor                      # (2,3) fp32
    grid: tuple                             # (N_r, N_theta, N_phi)
    r0: float = 0.03
    near: float = 0.01
    far: float = 15.0
    density_shift: float = -8.0
    distance_scale: float = 25.0
    n_coarse: int = 128
    n_fine: int = 128
    use_coarse_sample: bool = True
    resampling: bool = True
    fea2dense: str = "softplus"
    shading: str = "MLP_Fea"                # MLP_Fea | MLP | RGB | SH
    view_pe: int = 2
    fea_pe: int = 2
    app_dim: int = 27
    exp_sampling: bool = True               # False: uniform march of TensorBase.sample_ray (tensorBase.py:308-327)
    interval_th: bool = True                # False: the plain exponential ladders of a run without --interval_th (opt.py:190)
    step_ratio: float = 0.5
    extras: dict = field(default_factory=dict)


def yinyang_resolution(n_voxels: float):
    """coordinates.py:507-520 (YinYangSphericalCoords.N_to_reso)."""
    n_r = int(pow(n_voxels, 1 / 3) / 2)
    n_t = int(n_r * 2 * math.sqrt(3) / 3)
    n_p = n_t * 3
    n_r += n_r % 2
    n_t += n_t % 2
    n_p += n_p % 2
    return [n_r, n_t, n_p]


def scene_center(aabb: torch.Tensor) -> torch.Tensor:
    """coordinates.py:77 — center = aabb.sum(0) / 2."""
    return aabb.sum(0).div(2)


def max_corner_radius(aabb: torch.Tensor) -> torch.Tensor:
    """coordinates.py:187-204 (_get_max_r): largest centre→AABB-corner distance, fp32 0-dim tensor."""
    lo, hi = aabb.tolist()
    corners = torch.tensor([[lo[b] if (i >> b) & 1 else hi[b] for b in range(3)] for i in range(8)],
                           dtype=torch.float32)
    return (corners - scene_center(aabb)).pow(2).sum(1).sqrt().amax()


# --------------------------------------------------------------------------------------------------
# A1 / A4: exponential schedules with the "first K intervals forced to r0" fix-up
# --------------------------------------------------------------------------------------------------
def _exp_ladder(r0, ratio, idx: torch.Tensor) -> torch.Tensor:
    """extra/test_exp_r.py:10-15 (index2r): r_0 = 0, r_i = r0 * ratio**(i-1), fp32."""
    out = torch.zeros(idx.shape, dtype=torch.float32)
    pos = idx > 0
    out[pos] = r0 * ratio ** (idx[pos] - 1)
    return out


def _force_linear_prefix(r: torch.Tensor, r0: float) -> torch.Tensor:
    """EgoNeRF.py:72-76 and coordinates.py:120-124: intervals <= r0 become exactly r0, the rest shift."""
    iv = r[1:] - r[:-1]
    cum = torch.cumsum(iv, dim=0)
    k = (iv <= r0).sum()
    r = r.clone()
    r[:k + 1] = torch.arange(k + 1) * r0
    r[k + 1:] = r[k + 1:] + r0 * k - cum[k - 1]
    return r


def sample_schedule(near: float, far: float, r0: float, n: int) -> torch.Tensor:
    """EgoNeRF.sample_ray_exp, interval_th branch (EgoNeRF.py:68-76).  Returns r (n,), z = near + r."""
    idx = torch.arange(n).float()
    ratio = math.exp(math.log((far - near) / r0) / (n - 1))
    return _force_linear_prefix(_exp_ladder(r0, ratio, idx), r0)


def r_reference_grid(far_r: torch.Tensor, r0: float, n_r: int) -> torch.Tensor:
    """GenericSphericalCoords.normalize_r, interval_th branch (coordinates.py:112-124): N_r+1 knots."""
    ratio = pow(far_r / r0, 1 / (n_r - 1))          # fp32 0-dim tensor, like `self.far[0] / r0`
    return _force_linear_prefix(_exp_ladder(r0, ratio, torch.arange(n_r + 1)), r0)


def march_step_size(aabb: torch.Tensor, grid, step_ratio: float) -> torch.Tensor:
    """TensorBase.update_stepSize (tensorBase.py:206-213): mean(aabbSize / (gridSize - 1)) * step_ratio, fp32 0-dim."""
    units = (aabb[1] - aabb[0]) / (torch.tensor(list(grid), dtype=torch.int64) - 1)
    return torch.mean(units) * step_ratio


def uniform_march(o, d, aabb, near, far, step, n, u=None):
    """TensorBase.sample_ray (tensorBase.py:308-327): t_min = entry into the AABB clamped to [near, far];
    z_j = t_min + step * (j [+ U])."""
    vec = torch.where(d == 0, torch.full_like(d, 1e-6), d)
    rate_a = (aabb[1] - o) / vec
    rate_b = (aabb[0] - o) / vec
    t_min = torch.minimum(rate_a, rate_b).amax(-1).clamp(min=near, max=far)
    rng = torch.arange(n)[None].float()
    if u is not None:
        rng = rng.repeat(o.shape[0], 1) + u
    return t_min[..., None] + step * rng


def plain_sample_schedule(near: float, far: float, n: int, u: Optional[torch.Tensor] = None) -> torch.Tensor:
    """EgoNeRF.sample_ray_exp without interval_th (EgoNeRF.py:59-66): radii (1|N, n) WITHOUT near.  ratio = 1 + (pi/2)/n,
    r0' = (far-near)(ratio-1)/(ratio^n - 1); r_j = r0' * sum_{i<j} ratio^(i [+ u_i]) -- the jitter sits in the exponent.
    The reference forms the exclusive sums as a product with a strictly-lower-triangular ones matrix, transposed."""
    ratio = 1 + (PI / 2.) / n
    r0 = (far - near) * (ratio - 1) / (pow(ratio, n) - 1)
    e = torch.arange(n)[None].float()
    if u is not None:
        e = e.repeat(u.shape[0], 1) + u
    return torch.pow(ratio, e) @ torch.tril(torch.ones(n, n), diagonal=-1).T * r0


def jitter_schedule(r: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """EgoNeRF.py:77-81 (train): r + interval * U, last interval repeated.  r (n,), u (N,n)."""
    r = r.repeat(u.shape[0], 1)
    iv = r[:, 1:] - r[:, :-1]
    iv = torch.cat([iv, iv[:, -1:]], dim=-1)
    return r + iv * u


# --------------------------------------------------------------------------------------------------
# A3 / A4: cartesian -> Yin-Yang -> normalised grid coordinates
# --------------------------------------------------------------------------------------------------
def cart_to_yinyang(p: torch.Tensor, center: torch.Tensor):
    """YinYangSphericalCoords.from_cartesian (coordinates.py:468-498).
    Returns r, a (polar angle in the active grid), b (azimuth in the active grid), is_yang (bool)."""
    q = p - center
    r = q.pow(2).sum(-1).sqrt()
    th_n = torch.acos(q[..., 2] / r).nan_to_num_()
    ph_n = torch.atan2(q[..., 1], q[..., 0])
    yin = (PI / 4 <= th_n) & (th_n <= 3 * PI / 4) & (-3 * PI / 4 <= ph_n) & (ph_n <= 3 * PI / 4)
    th_e = torch.acos(q[..., 1] / r).nan_to_num_()
    ph_e = torch.atan2(q[..., 2], -q[..., 0])
    a = torch.where(yin, th_n, th_e)
    b = torch.where(yin, ph_n, ph_e)
    return r, a, b, ~yin, (th_n, ph_n)


def normalize_angles(a: torch.Tensor, b: torch.Tensor):
    """coordinates.py:458-459,500-505: affine map of [pi/4,3pi/4] x [-3pi/4,3pi/4] to [-1,1], fp32 constants."""
    near = torch.tensor([PI / 4, -3 * PI / 4], dtype=torch.float32)
    far = torch.tensor([3 * PI / 4, 3 * PI / 4], dtype=torch.float32)
    inv = 1.0 / (far - near)
    return (a - near[0]) * inv[0] * 2 - 1, (b - near[1]) * inv[1] * 2 - 1


def normalize_radius(r: torch.Tensor, knots: torch.Tensor) -> torch.Tensor:
    """coordinates.py:125-131,156 + 456: searchsorted(right) on the knot ladder, /N_r, *2-1."""
    n_r = knots.shape[0] - 1
    hi = torch.clamp(torch.searchsorted(knots, r.contiguous(), side="right"), 1, n_r)
    lo = hi - 1
    frac = (r - knots[lo]) / (knots[hi] - knots[lo])
    return (lo + frac) / n_r * 2 - 1


def normalize_radius_plain(r: torch.Tensor, far_r: torch.Tensor, r0: float, n_r: int, downsample=None) -> torch.Tensor:
    """GenericSphericalCoords.normalize_r without interval_th (coordinates.py:132-156) + the *2-1 of normalize_coord:
    k = trunc(log(r/r0)/log(ratio)); r < r0 -> r/r0, else 1 + k + (r - r0 ratio^k)/(r0 ratio^(k+1) - r0 ratio^k); / N_r.
    `downsample` halves N_r and recomputes the ratio (the coarse pass, EgoNeRF.py:523)."""
    if downsample is not None:
        n_r = n_r // downsample
    ratio = pow(far_r / r0, 1 / (n_r - 1))                  # fp32 0-dim tensor (coordinates.py:139,215)
    k = (torch.log(r / r0) / math.log(ratio)).to(torch.int32)
    below = r < r0
    lo = r0 * torch.pow(ratio, k)
    hi = r0 * torch.pow(ratio, k + 1)
    lo[below], hi[below] = 0, r0
    idx = torch.where(below, r / r0, 1 + k + (r - lo) / (hi - lo))
    return idx / n_r * 2 - 1


# --------------------------------------------------------------------------------------------------
# A5: factor gather with explicit taps (semantics of F.grid_sample bilinear/zeros/align_corners=True)
# --------------------------------------------------------------------------------------------------
def _tap2d(img: torch.Tensor, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """img (C,H,W); x->W, y->H in [-1,1]; returns (M,C).  Out-of-range taps contribute zero."""
    C, H, W = img.shape
    ix = ((x + 1) / 2) * (W - 1)
    iy = ((y + 1) / 2) * (H - 1)
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    x1, y1 = x0 + 1, y0 + 1
    w_nw = (x1 - ix) * (y1 - iy)
    w_ne = (ix - x0) * (y1 - iy)
    w_sw = (x1 - ix) * (iy - y0)
    w_se = (ix - x0) * (iy - y0)
    flat = img.reshape(C, H * W).t()                       # (H*W, C)

    def fetch(xx, yy):
        ok = (xx >= 0) & (xx <= W - 1) & (yy >= 0) & (yy <= H - 1)
        lin = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)).long()
        return flat[lin] * ok[:, None].to(flat.dtype)

    return (fetch(x0, y0) * w_nw[:, None] + fetch(x1, y0) * w_ne[:, None]
            + fetch(x0, y1) * w_sw[:, None] + fetch(x1, y1) * w_se[:, None])


def _tap1d(line: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """line (C,L); the reference samples an (L x 1) image at x = 0 (EgoNeRF.py:311-313)."""
    C, L = line.shape
    it = ((t + 1) / 2) * (L - 1)
    t0 = torch.floor(it)
    t1 = t0 + 1
    flat = line.t()

    def fetch(tt):
        ok = (tt >= 0) & (tt <= L - 1)
        return flat[tt.clamp(0, L - 1).long()] * ok[:, None].to(flat.dtype)

    return fetch(t0) * (t1 - it)[:, None] + fetch(t1) * (it - t0)[:, None]


MAT_MODE = ((0, 1), (0, 2), (1, 2))      # EgoNeRF.py:30-33
VEC_MODE = (2, 1, 0)


def avg_pool_factors(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """EgoNeRF.update_coarse_sigma_grid (EgoNeRF.py:124-133): 2x2 / 2 average pooling, floor sizes."""
    out = {}
    for h in ("yin", "yang"):
        for i in range(3):
            p = sd[f"density_plane_{h}.{i}"]
            l = sd[f"density_line_{h}.{i}"]
            out[f"coarse_plane_{h}.{i}"] = torch.nn.functional.avg_pool2d(p, 2, 2)
            out[f"coarse_line_{h}.{i}"] = torch.nn.functional.avg_pool1d(l.squeeze(-1), 2, 2).unsqueeze(-1)
    return out


# --------------------------------------------------------------------------------------------------
# f4: coarse-to-fine resampling of the factor tensors (train.py:371-377)
# --------------------------------------------------------------------------------------------------
def upsample_r_samples(far_r: torch.Tensor, r0: float, n_r_old: int, n_r_new: int) -> torch.Tensor:
    """GenericSphericalCoords.up_sampling_VM, interval_th branch (coordinates.py:238-246): the n_r_new radii of the target
    ladder (ratio recomputed for n_r_new, same r0, short intervals forced to r0) located on the CURRENT ladder, in [-1,1]."""
    ratio = pow(far_r / r0, 1 / (n_r_new - 1))
    target = _force_linear_prefix(_exp_ladder(r0, ratio, torch.arange(n_r_new)), r0)
    return normalize_radius(target, r_reference_grid(far_r, r0, n_r_old))


def _interp_align_corners(img: torch.Tensor, h2: int, w2: int) -> torch.Tensor:
    """F.interpolate(mode='bilinear', align_corners=True) of a (C,H,W) image (Coordinates.up_sampling_VM,
    coordinates.py:27-39): src = j * (in-1)/(out-1); index0 = trunc, index1 = index0 + 1 (clamped); lambda1 = src - index0."""
    C, H, W = img.shape

    def axis(n_in, n_out):
        if n_out == n_in:
            i0 = torch.arange(n_out)
            return i0, i0, torch.zeros(n_out)
        scale = (torch.tensor(float(n_in - 1)) / (n_out - 1)) if n_out > 1 else torch.zeros(())
        src = torch.arange(n_out, dtype=torch.float32) * scale
        i0 = src.long().clamp(max=n_in - 1)
        i1 = i0 + (i0 < n_in - 1).long()
        return i0, i1, (src - i0).clamp(0, 1)

    y0, y1, ly = axis(H, h2)
    x0, x1, lx = axis(W, w2)
    top = img[:, y0][:, :, x0] * (1 - lx) + img[:, y0][:, :, x1] * lx
    bot = img[:, y1][:, :, x0] * (1 - lx) + img[:, y1][:, :, x1] * lx
    return top * (1 - ly)[None, :, None] + bot * ly[None, :, None]


def upsample_factor(w: torch.Tensor, res_target, ids, far_r: torch.Tensor, r0: float, n_r_old: int) -> torch.Tensor:
    """One (1,C,H,W) factor tensor -> res_target (GenericSphericalCoords.up_sampling_VM coordinates.py:226-266; planes are
    passed with ids = [m1, m0], lines with ids = [v], EgoNeRF.py:415-425).  Axes without r: F.interpolate.  With r:
    F.grid_sample on (r_samples x linspace(-1,1))."""
    img = w[0]
    C, H, W = img.shape
    if 0 not in ids:
        h2 = res_target[ids[0]]
        w2 = res_target[ids[1]] if len(ids) == 2 else 1
        return _interp_align_corners(img, h2, w2)[None]
    rs = upsample_r_samples(far_r, r0, n_r_old, res_target[0])
    if len(ids) == 1:                                        # (C, N_r, 1) line: x = -1 on a width-1 image, y = r_samples
        return _tap1d(img[:, :, 0], rs).t().reshape(1, C, -1, 1)
    other = 1 - ids.index(0)
    lin = torch.linspace(-1, 1, res_target[ids[other]])
    if ids.index(0) == 1:                                    # (C, other, r): x = r, y = other
        xx, yy = torch.meshgrid(rs, lin, indexing="xy")      # (n_other, n_r)
    else:                                                    # (C, r, other)
        xx, yy = torch.meshgrid(lin, rs, indexing="xy")
    out = _tap2d(img, xx.reshape(-1), yy.reshape(-1))        # (h2*w2, C)
    return out.t().reshape(1, C, xx.shape[0], xx.shape[1])


def upsample_factors(sd: Dict[str, torch.Tensor], res_target, far_r: torch.Tensor, r0: float, n_r_old: int):
    """EgoNeRF.upsample_volume_grid (EgoNeRF.py:427-436) on a state dict: all 24 factor tensors, the rest copied."""
    out = dict(sd)
    for kind in ("density", "app"):
        for h in ("yin", "yang"):
            for i in range(3):
                m0, m1 = MAT_MODE[i]
                out[f"{kind}_plane_{h}.{i}"] = upsample_factor(sd[f"{kind}_plane_{h}.{i}"], res_target, [m1, m0], far_r, r0, n_r_old)
                out[f"{kind}_line_{h}.{i}"] = upsample_factor(sd[f"{kind}_line_{h}.{i}"], res_target, [VEC_MODE[i]], far_r, r0, n_r_old)
    return out


def _products(sd, plane_key, line_key, coords, is_yang):
    """Returns list over i<3 of (M, C_i) plane*line products for the active hemisphere of every sample."""
    M = coords.shape[0]
    outs = []
    for i in range(3):
        m0, m1 = MAT_MODE[i]
        v = VEC_MODE[i]
        C = sd[f"{plane_key}_yin.{i}"].shape[1]
        res = torch.zeros(M, C, dtype=torch.float32)
        for h, sel in (("yin", ~is_yang), ("yang", is_yang)):
            if sel.any():
                c = coords[sel]
                P = _tap2d(sd[f"{plane_key}_{h}.{i}"][0], c[:, m0], c[:, m1])
                L = _tap1d(sd[f"{line_key}_{h}.{i}"][0, :, :, 0], c[:, v])
                res[sel] = P * L
        outs.append(res)
    return outs


def density_feature(sd, coords, is_yang, coarse=False):
    """EgoNeRF.compute_densityfeature (EgoNeRF.py:291-347) / compute_coarse_densityfeature (:232-289)."""
    pk, lk = ("coarse_plane", "coarse_line") if coarse else ("density_plane", "density_line")
    f = torch.zeros(coords.shape[0], dtype=torch.float32)
    for prod in _products(sd, pk, lk, coords, is_yang):
        f = f + torch.relu(prod.sum(-1))
    return f


def app_feature(sd, coords, is_yang):
    """EgoNeRF.compute_appfeature (EgoNeRF.py:349-413): cat_i(P_i*L_i) @ basis_mat_{yin,yang}.T"""
    v = torch.cat(_products(sd, "app_plane", "app_line", coords, is_yang), dim=-1)   # (M, sum C)
    out = torch.zeros(coords.shape[0], sd["basis_mat_yin.weight"].shape[0], dtype=torch.float32)
    for h, sel in (("yin", ~is_yang), ("yang", is_yang)):
        if sel.any():
            out[sel] = v[sel] @ sd[f"basis_mat_{h}.weight"].t()
    return out


# --------------------------------------------------------------------------------------------------
# A6-A10
# --------------------------------------------------------------------------------------------------
def feature_to_density(f, shift, act="softplus"):
    """tensorBase.py:415-419."""
    return torch.nn.functional.softplus(f + shift) if act == "softplus" else torch.relu(f)


# --------------------------------------------------------------------------------------------------
# f4: occupancy mask (deprecated in the reference; only compute_alpha reads it)
# --------------------------------------------------------------------------------------------------
def sample_alpha_mask(vol_yin, vol_yang, coords7):
    """YinYangAlphaGridMask.sample_alpha (EgoNeRF.py:19-24): trilinear, zeros padding, align_corners, volumes (1,1,D,H,W)
    with x = r -> W."""
    out = torch.empty(coords7.shape[0])
    yin = coords7[:, 6] == 0
    gs = torch.nn.functional.grid_sample
    out[yin] = gs(vol_yin, coords7[yin][:, :3].view(1, -1, 1, 1, 3), align_corners=True).view(-1)
    out[~yin] = gs(vol_yang, coords7[~yin][:, 3:6].view(1, -1, 1, 1, 3), align_corners=True).view(-1)
    return out


def compute_alpha(sd, coords7, step, shift, mask=None, act="softplus"):
    """TensorBase.compute_alpha (tensorBase.py:421-436) at normalised 7-coords; `mask` = (vol_yin, vol_yang) or None."""
    keep = torch.ones(coords7.shape[0], dtype=torch.bool) if mask is None else sample_alpha_mask(mask[0], mask[1], coords7) > 0
    sigma = torch.zeros(coords7.shape[0])
    if keep.any():
        c = coords7[keep]
        is_yang = c[:, 6] != 0
        c3 = torch.where(is_yang[:, None], c[:, 3:6], c[:, 0:3])
        sigma[keep] = feature_to_density(density_feature(sd, c3, is_yang), shift, act)
    return 1 - torch.exp(-sigma * step)


def dense_alpha(sd, grid, step, shift):
    """EgoNeRF.getDenseAlpha (EgoNeRF.py:438-465): (g0,g1,g2) alpha lattices of the Yin and the Yang hemisphere."""
    lin = [torch.linspace(0, 1, g) * 2 - 1 for g in grid]
    pts = torch.stack(torch.meshgrid(*lin, indexing="ij"), -1).reshape(-1, 3)
    out = []
    for h in (0, 1):
        c7 = torch.zeros(pts.shape[0], 7)
        c7[:, 3 * h:3 * h + 3] = pts
        c7[:, 6] = h
        out.append(compute_alpha(sd, c7, step, shift).view(*grid))
    return out


def alpha_mask_volumes(alpha_yin, alpha_yang, thres):
    """EgoNeRF.updateAlphaMask (EgoNeRF.py:471-486): clamp, (r,theta,phi) -> (phi,theta,r), 3^3 max-pool, binarise."""
    vols = []
    for a in (alpha_yin, alpha_yang):
        a = a.clamp(0, 1).transpose(0, 2).contiguous()[None, None]
        a = torch.nn.functional.max_pool3d(a, kernel_size=3, padding=1, stride=1)
        vols.append((a >= thres).float())
    return vols


def alpha_composite_weights(sigma, dist):
    """raw2alpha (tensorBase.py:22-27)."""
    alpha = 1. - torch.exp(-sigma * dist)
    T = torch.cumprod(torch.cat([torch.ones(alpha.shape[0], 1), 1. - alpha + 1e-10], -1), -1)
    return alpha, alpha * T[:, :-1], T[:, -1:]


def inverse_cdf(bins, weights, u):
    """sample_pdf (dataLoader/ray_utils.py:156-187).  bins (N,B), weights (N,B-1), u (N,n)."""
    w = weights + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    u = u.contiguous()
    ind = torch.searchsorted(cdf, u, right=True)
    below = (ind - 1).clamp(min=0)
    above = ind.clamp(max=cdf.shape[-1] - 1)
    c0, c1 = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    b0, b1 = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    den = c1 - c0
    den = torch.where(den < 1e-5, torch.ones_like(den), den)
    return b0 + (u - c0) / den * (b1 - b0), den


def freq_encode(x, n_freq):
    """positional_encoding (tensorBase.py:14-19): [sin(x_j 2^f)], [cos(x_j 2^f)], index j*F+f."""
    bands = 2 ** torch.arange(n_freq).float()
    p = (x[..., None] * bands).reshape(x.shape[:-1] + (n_freq * x.shape[-1],))
    return torch.cat([torch.sin(p), torch.cos(p)], dim=-1)


SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)


def sh_deg2(d):
    """eval_sh_bases(2, dirs) (models/sh.py:87-116)."""
    x, y, z = d.unbind(-1)
    return torch.stack([
        torch.full_like(x, SH_C0), -SH_C1 * y, SH_C1 * z, -SH_C1 * x,
        SH_C2[0] * (x * y), SH_C2[1] * (y * z), SH_C2[2] * (2.0 * z * z - x * x - y * y),
        SH_C2[3] * (x * z), SH_C2[4] * (x * x - y * y)], dim=-1)


def decode_color(sd, cfg: OracleCfg, feat, dirs):
    """renderModule: MLPRender_Fea (tensorBase.py:54-78), MLPRender (:107-129), RGBRender (:37-39),
    SHRender (:30-34, on flattened inputs — the reference's own EgoNeRF.forward crashes in SH mode)."""
    if cfg.shading == "RGB":
        return feat
    if cfg.shading == "SH":
        Y = sh_deg2(dirs)[:, None]                                   # (M,1,9)
        return torch.relu(torch.sum(Y * feat.view(-1, 3, 9), dim=-1) + 0.5)
    x = [feat, dirs]
    if cfg.shading == "MLP_Fea" and cfg.fea_pe > 0:
        x.append(freq_encode(feat, cfg.fea_pe))
    if cfg.view_pe > 0:
        x.append(freq_encode(dirs, cfg.view_pe))
    h = torch.cat(x, dim=-1)
    h = torch.relu(h @ sd["renderModule.mlp.0.weight"].t() + sd["renderModule.mlp.0.bias"])
    h = torch.relu(h @ sd["renderModule.mlp.2.weight"].t() + sd["renderModule.mlp.2.bias"])
    h = h @ sd["renderModule.mlp.4.weight"].t() + sd["renderModule.mlp.4.bias"]
    return torch.sigmoid(h)


def envmap_radiance(emission, dirs):
    """EnvironmentMap.get_radiance (models/envmap.py:6-34).  emission (3,2h,h)."""
    d = torch.nn.functional.normalize(dirs, dim=-1)
    u = (d[:, 2] + 1) * 0.5
    v = (torch.atan2(d[:, 1], d[:, 0]) + PI) / (2 * PI)
    return torch.sigmoid(_tap2d(emission, 2 * u - 1, 2 * v - 1))


# --------------------------------------------------------------------------------------------------
# the whole path
# --------------------------------------------------------------------------------------------------
def _coords(p, center, knots):
    """`knots`: the interval_th ladder, or a callable r -> normalised r (plain ladders)."""
    r, a, b, is_yang, (th_n, ph_n) = cart_to_yinyang(p, center)
    an, bn = normalize_angles(a, b)
    rn = knots(r) if callable(knots) else normalize_radius(r, knots)
    margin = torch.stack([(th_n - PI / 4).abs(), (th_n - 3 * PI / 4).abs(),
                          (ph_n + 3 * PI / 4).abs(), (ph_n - 3 * PI / 4).abs()], -1).amin(-1)
    return torch.stack([rn, an, bn], -1), is_yang, margin


def render(sd: Dict[str, torch.Tensor], cfg: OracleCfg, rays: torch.Tensor, is_train: bool = False,
           u_coarse: Optional[torch.Tensor] = None, u_fine: Optional[torch.Tensor] = None,
           emission: Optional[torch.Tensor] = None, want_aux: bool = False):
    """EgoNeRF.forward (EgoNeRF.py:491-602) with exp_sampling + interval_th, as every shipped config runs it.
    rays (N,6) = [o, d].  Train mode needs the two uniform draws the reference takes from its RNGs:
    u_coarse (N,n_coarse) (EgoNeRF.py:81) and u_fine (N,n_fine) (ray_utils.py:169)."""
    N = rays.shape[0]
    o, d = rays[:, :3], rays[:, 3:6]
    center = scene_center(cfg.aabb)
    far_r = max_corner_radius(cfg.aabb)
    if cfg.interval_th:
        knots = knots_c = r_reference_grid(far_r, cfg.r0, cfg.grid[0])       # `downsample` is ignored (coordinates.py:112-117)
    else:
        def knots(r):
            return normalize_radius_plain(r, far_r, cfg.r0, cfg.grid[0])

        def knots_c(r):                                                        # normalize_coord(downsample=2), EgoNeRF.py:523
            return normalize_radius_plain(r, far_r, cfg.r0, cfg.grid[0], downsample=2)
    nc, nf = cfg.n_coarse, cfg.n_fine

    if cfg.exp_sampling and not cfg.interval_th:
        zc = cfg.near + plain_sample_schedule(cfg.near, cfg.far, nc, u_coarse if is_train else None)
        zc = zc if is_train else zc[0].repeat(N, 1)
        zq = zc
    elif cfg.exp_sampling:
        r = sample_schedule(cfg.near, cfg.far, cfg.r0, nc)
        if is_train:
            zc = cfg.near + jitter_schedule(r, u_coarse)
        else:
            zc = (cfg.near + r).repeat(N, 1)
        zq = zc
    else:
        # uniform march; the coarse POINTS use every ray's own depths, but in eval mode the depths handed on are those of
        # the first ray of the chunk (EgoNeRF.py:515-516: coarse_z_vals[0].repeat(N, 1)) — reproduced as is
        step = march_step_size(cfg.aabb, cfg.grid, cfg.step_ratio)
        zq = uniform_march(o, d, cfg.aabb, cfg.near, cfg.far, step, nc, u_coarse if is_train else None)
        zc = zq if is_train else zq[0].repeat(N, 1)
    dc = zc[:, 1:] - zc[:, :-1]
    dc = torch.cat([dc, dc[:, -1:]], -1)
    pc = o[:, None, :] + d[:, None, :] * zq[..., None]
    cc, yang_c, margin_c = _coords(pc.reshape(-1, 3), center, knots_c)
    aux = {}

    if cfg.resampling:
        pooled = avg_pool_factors(sd)
        sig_c = feature_to_density(density_feature(pooled, cc, yang_c, coarse=True), cfg.density_shift,
                                   cfg.fea2dense).view(N, nc)
        _, w_c, _ = alpha_composite_weights(sig_c, dc * cfg.distance_scale)
        mid = .5 * (zc[:, 1:] + zc[:, :-1])
        if not is_train:
            u_fine = torch.linspace(0., 1., steps=nf).expand(N, nf)
        z_new, den = inverse_cdf(mid, w_c[:, 1:-1], u_fine)
        z_new = z_new.detach()
        z = torch.sort(torch.cat([zc, z_new], -1) if cfg.use_coarse_sample else z_new, -1)[0]
        dist = z[:, 1:] - z[:, :-1]
        dist = torch.cat([dist, dist[:, -1:]], -1)
        pts = o[:, None, :] + d[:, None, :] * z[..., None]
        cf, yang_f, margin_f = _coords(pts.reshape(-1, 3), center, knots)
        aux.update(coarse_sigma=sig_c, coarse_weight=w_c, z_new=z_new, cdf_den=den)
    else:
        z, dist, cf, yang_f, margin_f = zc, dc, cc, yang_c, margin_c
    S = z.shape[1]

    sigma = feature_to_density(density_feature(sd, cf, yang_f), cfg.density_shift, cfg.fea2dense).view(N, S)
    alpha, w, bg = alpha_composite_weights(sigma, dist * cfg.distance_scale)
    feat = app_feature(sd, cf, yang_f)
    dirs = d[:, None, :].expand(N, S, 3).reshape(-1, 3)
    rgb_s = decode_color(sd, cfg, feat, dirs).view(N, S, 3)

    acc = torch.sum(w, -1)
    rgb = torch.sum(w[..., None] * rgb_s, -2)
    bg_map = env = None
    if emission is not None:
        alpha = torch.cat((alpha, torch.ones_like(alpha[..., :1])), dim=-1)
        env = envmap_radiance(emission, d)
        bg_map = bg * env
        rgb = rgb + bg_map
    rgb = rgb.clamp(0, 1)
    with torch.no_grad():
        depth = torch.sum(w * z, -1) + (1. - acc) * rays[..., -1]
    if want_aux:
        aux.update(z=z, sigma=sigma, weight=w, rgb_samples=rgb_s, feat=feat.view(N, S, -1),
                   coords=cf.view(N, S, 3), is_yang=yang_f.view(N, S),
                   margin=torch.minimum(margin_f.view(N, S).amin(-1),
                                        margin_c.view(N, nc).amin(-1)) if cfg.resampling
                   else margin_f.view(N, S).amin(-1))
        return (rgb, depth, bg_map, env, alpha), aux
    return rgb, depth, bg_map, env, alpha
