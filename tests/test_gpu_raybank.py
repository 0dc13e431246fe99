"""GPU: device-resident batch assembly: ERP ray generation from a pose and the on-device epoch sampler."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference_erp(H, W, c2w):
    """get_ray_directions_360 + normalisation + get_rays, restated with torch on the CPU
    (dataLoader/ray_utils.py:24-40,85-113; dataset_omniblender.py:42-43)."""
    i = torch.tile(torch.arange(W), (H, 1)) + 0.5
    j = torch.tile(torch.arange(H), (W, 1)).T + 0.5
    phi = (1 - 2 * i / W) * np.pi
    theta = (1 - 2 * j / H) * np.pi / 2
    d = torch.stack([-torch.cos(theta) * torch.sin(phi), torch.sin(theta), -torch.cos(theta) * torch.cos(phi)], -1)
    d = d / torch.norm(d, dim=-1, keepdim=True)
    rd = (d @ c2w[:3, :3].T).view(-1, 3)
    ro = c2w[:3, 3].expand(rd.shape)
    return torch.cat([ro, rd], 1)


def test_erp_rays_match_reference_formula():
    from egonerf_b200.raybank import erp_rays
    g = torch.Generator().manual_seed(3)
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
    c2w = torch.cat([q, torch.tensor([[0.3], [-0.1], [0.2]])], 1)
    H, W = 64, 128
    ref = _reference_erp(H, W, c2w)
    out = erp_rays(H, W, c2w).cpu()
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() <= 2e-6
    tile = erp_rays(H, W, c2w, rows=(10, 23)).cpu()
    assert torch.equal(tile, out[10 * W:23 * W])
    assert abs(float(out[:, 3:].norm(dim=-1).mean()) - 1.0) < 1e-6


def test_erp_rays_match_the_reference_fixture():
    """tests/golden/erp_rays_ref.npz: `get_ray_directions_360` + normalisation + `get_rays` of the UNMODIFIED reference
    (dataLoader/ray_utils.py:24-40,85-113; dataset_omniblender.py:42-43), frozen by oracle/make_golden.py."""
    import os
    from egonerf_b200.raybank import erp_rays
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "erp_rays_ref.npz"))
    H, W = int(g["H"]), int(g["W"])
    out = erp_rays(H, W, torch.from_numpy(g["c2w"])).cpu().numpy()
    assert out.shape == g["rays"].shape
    assert np.abs(out - g["rays"]).max() <= 2e-6


def test_raybank_epoch_semantics():
    from egonerf_b200.raybank import RayBank
    n, batch = 1000, 96
    rays = torch.arange(n, dtype=torch.float32)[:, None].repeat(1, 6)
    rgbs = torch.arange(n, dtype=torch.float32)[:, None].repeat(1, 3)
    bank = RayBank(rays, rgbs, batch, device="cuda:0", seed=1)
    seen = []
    for _ in range(n // batch):                      # one full permutation: disjoint batches
        r, c = bank.next_batch()
        assert r.shape == (batch, 6) and torch.equal(r[:, 0], c[:, 0])
        seen.append(r[:, 0].long().cpu())
    ids = torch.cat(seen)
    assert ids.unique().numel() == ids.numel()
    first_epoch = bank.ids.clone()
    bank.next_batch()                                # fewer than `batch` unseen rays remain -> new permutation
    assert not torch.equal(first_epoch, bank.ids)
    # ray-sharded banks are disjoint and cover everything
    parts = [RayBank(rays, rgbs, batch, device="cuda:0", rank=r, world=3) for r in range(3)]
    allr = torch.cat([p.rays[:, 0] for p in parts]).cpu()
    assert torch.equal(allr, torch.arange(n, dtype=torch.float32))
