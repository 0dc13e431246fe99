"""Host-side mirror of the reference's Yin-Yang coordinate object (models/coordinates.py:432-520 on top of
GenericSphericalCoords :73-266).  It holds the scalars and the two exponential ladders the kernels need; the
per-sample arithmetic (from_cartesian / normalize_coord) runs in libegn_b200 (`egn_yinyang_coords`).

Only what the EgoNeRF path uses is mirrored: `exp_r=True`, with `interval_th=True` (what every shipped config selects,
configs/EgoNeRF/common.txt:2-3,14) or without it (the CLI default, opt.py:190).  The other eight coordinate systems of
the reference are out of scope.
"""
from __future__ import annotations

from math import exp, log, pi, sqrt

import torch

from .. import _lib


def exp_ladder(r0, ratio, index: torch.Tensor) -> torch.Tensor:
    """r_0 = 0, r_i = r0 * ratio**(i-1) in fp32 (extra/test_exp_r.py:10-15)."""
    r = torch.zeros(index.shape, dtype=torch.float32)
    nz = index > 0
    r[nz] = r0 * ratio ** (index[nz] - 1)
    return r


def clamp_short_intervals(r: torch.Tensor, r0: float) -> torch.Tensor:
    """Intervals not longer than r0 become exactly r0 and the tail is shifted to stay continuous
    (EgoNeRF.py:72-76, coordinates.py:120-124)."""
    step = r[1:] - r[:-1]
    n_short = int((step <= r0).sum())
    run = torch.cumsum(step, dim=0)
    out = r.clone()
    out[:n_short + 1] = torch.arange(n_short + 1) * r0
    out[n_short + 1:] = r[n_short + 1:] + r0 * (step <= r0).sum() - run[n_short - 1]
    return out


def sample_schedule(near: float, far: float, r0: float, n: int) -> torch.Tensor:
    """Radii of the n coarse samples (without `near`), EgoNeRF.sample_ray_exp interval_th branch (EgoNeRF.py:69-76)."""
    ratio = exp(log((far - near) / r0) / (n - 1))
    return clamp_short_intervals(exp_ladder(r0, ratio, torch.arange(n).float()), r0)


def plain_sample_schedule(near: float, far: float, n: int):
    """Without interval_th (EgoNeRF.py:59-66): r_j = r0' * sum_{i<j} ratio^i with ratio = 1 + (pi/2)/n and r0' chosen so that
    the n intervals span far - near.  Returns (radii (n,), ratio, r0'); the reference sums with a (1,n) x (n,n) product
    against a strictly-upper-triangular ones matrix, kept here so that the eval depths are the reference's bit for bit."""
    ratio = 1 + (pi / 2.) / n
    r0 = (far - near) * (ratio - 1) / (pow(ratio, n) - 1)
    terms = torch.pow(ratio, torch.arange(n)[None].float())
    before = torch.tril(torch.ones(n, n), diagonal=-1).T            # [i, j] = 1 for i < j (same operand layout as the reference)
    return (terms @ before * r0)[0], ratio, r0


class YinYangSphericalCoords:
    """[r_n, theta_n, phi_n, r_e, theta_e, phi_e, Y]: Y = 0 Yin grid, 1 Yang grid."""

    def __init__(self, device, aabb, exp_r=True, N_voxel=None, r0=None, interval_th=False):
        if not exp_r:
            raise NotImplementedError("egonerf_b200 implements the exponential-r Yin-Yang grid only "
                                      "(exp_sampling is set by every shipped EgoNeRF config)")
        self.device = device
        self.aabb = aabb.to(device)
        self.center = self.aabb.sum(0).div(2)
        self.exp_r = exp_r
        self.interval_th = interval_th
        self.update_aabb(aabb)
        self.set_resolution(self.N_to_reso(N_voxel, aabb), r0=r0)

    def __setstate__(self, state):
        """Unpickling: also accepts the attribute set of the REFERENCE's object (center, device, near, far, inv_diff, exp_r,
        interval_th, N_r, N_theta, N_phi, r0, ratio) — `EgoNeRF.save` pickles the coordinates object into `kwargs`
        (EgoNeRF.py:158-160, tensorBase.py:241-268) and `train.py:155-160` rebuilds the model from it."""
        self.__dict__.update(state)
        self.__dict__.setdefault("aabb", None)
        self._knots = None
        if not self.exp_r:
            raise NotImplementedError("egonerf_b200 implements the exponential-r Yin-Yang grid only")

    # ---- scalars (coordinates.py:187-204, 500-505) --------------------------------------------------
    def _get_max_r(self, aabb):
        lo, hi = aabb.tolist()
        corners = torch.tensor([[lo[b] if (i >> b) & 1 else hi[b] for b in range(3)] for i in range(8)],
                               dtype=torch.float32)
        return (corners - self.center.cpu()).pow(2).sum(1).sqrt().amax()

    def update_aabb(self, new_aabb):
        max_r = self._get_max_r(new_aabb.cpu())
        self.near = torch.tensor([0, pi / 4, -3 * pi / 4, 0, pi / 4, -3 * pi / 4], dtype=torch.float32, device=self.device)
        self.far = torch.tensor([max_r, 3 * pi / 4, 3 * pi / 4, max_r, 3 * pi / 4, 3 * pi / 4], dtype=torch.float32,
                                device=self.device)
        self.inv_diff = 1.0 / (self.far - self.near)

    def N_to_reso(self, n_voxels, bbox=None):
        n_r = int(pow(n_voxels, 1 / 3) / 2)
        n_t = int(n_r * 2 * sqrt(3) / 3)
        n_p = n_t * 3
        return [n_r + n_r % 2, n_t + n_t % 2, n_p + n_p % 2]

    def set_resolution(self, resolution, r0=None):
        self.N_r, self.N_theta, self.N_phi = resolution
        self.r0 = r0 if r0 is not None else 0.05
        self.ratio = pow(self.far[0].cpu() / self.r0, 1 / (self.N_r - 1))
        self._knots = None

    # ---- ladders ------------------------------------------------------------------------------------
    def r_knots(self, downsample=None) -> torch.Tensor:
        """interval_th: the N_r + 1 knot radii normalize_r rebuilds on every call (coordinates.py:118-124; `downsample` is
        ignored, :112-117).  Otherwise: the knots r0 * ratio^k behind the closed form of :132-155, [0, r0, r0*ratio, ...],
        for N_r // downsample cells with the ratio recomputed (:137-139), two knots past the grid.  fp32, CPU."""
        if self.interval_th:
            if self._knots is None:
                ratio = pow(self.far[0].cpu() / self.r0, 1 / (self.N_r - 1))
                self._knots = clamp_short_intervals(exp_ladder(self.r0, ratio, torch.arange(self.N_r + 1)), self.r0)
            return self._knots
        n_r = self.N_r if downsample is None else self.N_r // downsample
        ratio = self.ratio if downsample is None else pow(self.far[0].cpu() / self.r0, 1 / (n_r - 1))
        knots = torch.zeros(n_r + 3)
        knots[1:] = self.r0 * torch.pow(ratio, torch.arange(n_r + 2, dtype=torch.int32))
        return knots

    def normalize_r(self, r: torch.Tensor) -> torch.Tensor:
        """Host restatement of GenericSphericalCoords.normalize_r, interval_th branch (coordinates.py:112-131,156) for
        the short ladders of `up_sampling_positions` (a few hundred radii): knot index + linear fraction, /N_r."""
        g = self.r_knots()                  # plain ladders: in + frac == 1 + k + lin of coordinates.py:141-155
        hi = torch.clamp(torch.searchsorted(g, r.contiguous(), side='right'), 1, g.shape[0] - 1)
        lo = hi - 1
        return (lo + (r - g[lo]) / (g[hi] - g[lo])) / self.N_r

    # ---- coarse-to-fine (coordinates.py:27-39, 226-266) ---------------------------------------------
    def up_sampling_positions(self, axis: int, n_in: int, n_out: int, beside_r: bool = False) -> torch.Tensor:
        """Source position (texel units, fp32, CPU) of each of the `n_out` samples along grid axis `axis` (0 = r).
        r: the target ladder (same r0, ratio recomputed for n_out knots, short intervals clamped) normalised on the
        CURRENT ladder and mapped through grid_sample's align_corners rule (coordinates.py:238-246,258-264);
        theta / phi: F.interpolate(align_corners=True), src = j * (n_in - 1) / (n_out - 1) (coordinates.py:38-39) -- or,
        for the angular axis of a plane that also has an r axis (`beside_r`), linspace(-1, 1) through grid_sample's rule
        (coordinates.py:250-264); the two differ by rounding only."""
        if axis == 0:
            ratio = pow(self.far[0].cpu() / self.r0, 1 / (n_out - 1))
            if self.interval_th:
                target = clamp_short_intervals(exp_ladder(self.r0, ratio, torch.arange(n_out)), self.r0)
            else:                                                    # coordinates.py:247-250
                target = torch.zeros(n_out)
                target[1:] = self.r0 * torch.pow(ratio, torch.arange(n_out - 1))
            r_samples = self.normalize_r(target) * 2 - 1
            return ((r_samples + 1) / 2) * (n_in - 1)
        if beside_r:
            return ((torch.linspace(-1, 1, n_out) + 1) / 2) * (n_in - 1)
        scale = torch.tensor(float(n_in - 1), dtype=torch.float32) / (n_out - 1) if n_out > 1 else torch.zeros(())
        return torch.arange(n_out, dtype=torch.float32) * scale

    def up_sampling_VM(self, weights: torch.Tensor, res_target, ids):
        """coordinates.py:226-266 (r axis) / :27-39 (angular axes) for one (1, C, res[ids[0]], res[ids[1]] or 1) factor
        tensor; the resampling itself runs in libegn_b200 (`egn_resample_factor`)."""
        assert len(ids) in (1, 2), 'ids should be 1 or 2!'
        if not weights.is_cuda:
            raise RuntimeError("egonerf_b200 runs on CUDA tensors only (no CPU fallback)")
        lib = _lib.load()
        src = weights.detach().contiguous().float()
        _, C, H, W = src.shape
        H2 = int(res_target[ids[0]])
        has_r = 0 in ids
        ypos = self.up_sampling_positions(ids[0], H, H2, has_r)
        if len(ids) == 2:
            W2 = int(res_target[ids[1]])
            xpos = self.up_sampling_positions(ids[1], W, W2, has_r)
        else:
            W2, xpos = 1, torch.zeros(1)
        ypos, xpos = ypos.to(src.device).contiguous(), xpos.to(src.device).contiguous()
        dst = torch.empty(1, C, H2, W2, device=src.device, dtype=torch.float32)
        _lib.check(lib.egn_resample_factor(src.data_ptr(), C, H, W, ypos.data_ptr(), H2, xpos.data_ptr(), W2, dst.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream))
        return torch.nn.Parameter(dst)

    # ---- operators (run in libegn_b200) -------------------------------------------------------------
    def _cfg(self):
        cfg = _lib.EgnConfig()
        cfg.grid[:] = [self.N_r, self.N_theta, self.N_phi]
        cfg.c_sigma, cfg.c_app, cfg.app_dim, cfg.shading, cfg.feature_c = 16, 48, 27, 2, 128
        cfg.app_dim = 3
        cfg.exp_sampling = 1
        c = self.center.cpu().tolist()
        cfg.center[:] = c
        near, inv = self.near.cpu(), self.inv_diff.cpu()
        cfg.ang_near[:] = [float(near[1]), float(near[2])]
        cfg.ang_inv[:] = [float(inv[1]), float(inv[2])]
        self._knots_dev = self.r_knots().to(self.device).contiguous()
        cfg.r_knots = self._knots_dev.data_ptr()
        cfg.plain_ladders = int(not self.interval_th)
        return cfg

    def cart_to_normalized(self, xyz: torch.Tensor) -> torch.Tensor:
        """normalize_coord(from_cartesian(xyz)) (coordinates.py:442-498) for (..., 3) device points."""
        lib = _lib.load()
        if not xyz.is_cuda:
            raise RuntimeError("egonerf_b200 runs on CUDA tensors only (no CPU fallback)")
        flat = xyz.reshape(-1, 3).contiguous().float()
        out = torch.empty(flat.shape[0], 7, device=xyz.device, dtype=torch.float32)
        cfg = self._cfg()
        _lib.check(lib.egn_yinyang_coords(cfg, flat.data_ptr(), flat.shape[0], out.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream))
        return out.view(*xyz.shape[:-1], 7)


coordinates_dict = {"yinyang": YinYangSphericalCoords}
