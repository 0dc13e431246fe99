"""CPU: the oracle (oracle/egn_oracle.py) held to the frozen outputs of the UNMODIFIED reference (tests/golden/*.npz,
made by oracle/make_golden.py).  This is the parity pin of the oracle — the reference ships no tests of its own.

Tolerances: the oracle is an independent fp32 restatement (explicit taps instead of F.grid_sample, gathers instead of
boolean-mask compaction), so sums are associated differently: 2e-6 on O(1) quantities, 1e-5 on rgb after 256-term sums.
"""
import os

import numpy as np
import pytest
import torch

from oracle import egn_oracle as O
from tests.helpers import GOLDEN, RENDER_CASES, TINY, T, checksum, load_golden, oracle_cfg, scene_for, stable_rays

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FAST = [n for n in RENDER_CASES if "tiny" in n]
SLOW = [n for n in RENDER_CASES if "tiny" not in n]


@pytest.mark.parametrize("tag", ["indoor300", "indoor128", "outdoor300"])
def test_schedule_and_coordinates(tag):
    g = load_golden("kat_coords_" + tag)
    aabb, r0 = T(g["aabb"]), float(g["r0"])
    near, far = [float(x) for x in g["near_far"]]
    grid = g["grid"].tolist()
    nvox = {"indoor300": 27e6, "indoor128": 128 ** 3, "outdoor300": 27e6}[tag]
    assert O.yinyang_resolution(nvox) == grid
    far_r = O.max_corner_radius(aabb)
    assert float(far_r) == float(g["far_r"])
    z = near + O.sample_schedule(near, far, r0, 128)
    assert np.array_equal(z.numpy(), g["z_coarse"]), "sample schedule must be bit-exact (EgoNeRF.py:68-76)"
    knots = O.r_reference_grid(far_r, r0, grid[0])
    pts = T(g["points"])
    r, a, b, is_yang, _ = O.cart_to_yinyang(pts, O.scene_center(aabb))
    ref_u, ref_n = g["unnormalized"], g["normalized"]
    assert np.array_equal(is_yang.numpy(), ref_u[:, 6] != 0)
    col = np.where(is_yang.numpy(), 3, 0)
    rows = np.arange(len(col))
    assert np.array_equal(r.numpy(), ref_u[rows, col])
    assert np.array_equal(a.numpy(), ref_u[rows, col + 1])
    assert np.array_equal(b.numpy(), ref_u[rows, col + 2])
    an, bn = O.normalize_angles(a, b)
    rn = O.normalize_radius(r, knots)
    assert np.abs(rn.numpy() - ref_n[rows, col]).max() <= 2e-7
    assert np.array_equal(an.numpy(), ref_n[rows, col + 1])
    assert np.array_equal(bn.numpy(), ref_n[rows, col + 2])


def test_known_answer_values():
    """SURVEY.md §8c KATs (aabb = ±15.5 cube, N_voxel 27e6, r0 .03, interval_th)."""
    aabb = torch.tensor([[-15.5] * 3, [15.5] * 3])
    far_r = O.max_corner_radius(aabb)
    assert abs(float(far_r) - 26.84678840637207) < 1e-6
    r = O.sample_schedule(0.01, 15., 0.03, 128)
    z = 0.01 + r
    assert np.allclose(z[:4].numpy(), [.01, .04, .07, .10], atol=1e-7)
    assert np.allclose(z[-4:].numpy(), [13.6022606, 14.2203226, 14.8693781, 15.5509796], atol=2e-6)
    knots = O.r_reference_grid(far_r, 0.03, 150)
    iv = knots[1:] - knots[:-1]
    assert int((torch.abs(iv - 0.03) < 1e-6).sum()) == 69
    c, yang, _ = O._coords(torch.tensor([[.3, -.2, .1], [-5., .5, 7.], [10., 10., 10.]]), O.scene_center(aabb), knots)
    assert np.allclose(c[0].numpy(), [-0.8337041, -0.3444746, -0.2495561], atol=2e-6) and not yang[0]
    assert np.allclose(c[1].numpy(), [0.6158372, -0.0739223, 0.4034245], atol=2e-6) and yang[1]
    assert np.allclose(c[2].numpy(), [0.8471348, -0.7836531, 0.3333334], atol=2e-6) and not yang[2]


def test_operators():
    g = load_golden("ops_small")
    scene = scene_for(dict(n_voxels=40 ** 3, seed=7))
    assert np.allclose(checksum(scene.state_dict), g["checksum"], rtol=1e-6)
    sd = dict(scene.state_dict)
    c7 = T(g["coords7"])
    yang = c7[:, 6] != 0
    c3 = torch.where(yang[:, None], c7[:, 3:6], c7[:, 0:3])
    assert np.abs(O.density_feature(sd, c3, yang).numpy() - g["sigma_feature"]).max() <= 1e-5
    sd.update(O.avg_pool_factors(sd))
    assert np.abs(O.density_feature(sd, c3, yang, coarse=True).numpy() - g["coarse_sigma_feature"]).max() <= 1e-5
    app = O.app_feature(sd, c3, yang)
    assert np.abs(app.numpy() - g["app_feature"]).max() <= 1e-5
    rgb = O.decode_color(sd, oracle_cfg(scene), T(g["app_feature"]), T(g["dirs"]))
    assert np.abs(rgb.numpy() - g["rgb"]).max() <= 2e-6


def test_composite_and_inverse_cdf():
    g = load_golden("composite_pdf")
    a, w, bg = O.alpha_composite_weights(T(g["sigma"]), T(g["dist"]))
    assert np.array_equal(a.numpy(), g["alpha"])
    assert np.abs(w.numpy() - g["weight"]).max() <= 1e-7 and np.abs(bg.numpy() - g["bg"]).max() <= 1e-7
    bins, wgt = T(g["bins"]), T(g["weight"])[:, 1:-1]
    ze, _ = O.inverse_cdf(bins, wgt, torch.linspace(0., 1., 128).expand(16, 128))
    zt, _ = O.inverse_cdf(bins, wgt, T(g["u_train"]))
    assert np.abs(ze.numpy() - g["fine_eval"]).max() <= 1e-5
    assert np.abs(zt.numpy() - g["fine_train"]).max() <= 1e-5


def _run_case(name):
    skw, okw = RENDER_CASES[name]
    g = load_golden(name)
    scene = scene_for(skw)
    assert np.allclose(checksum(scene.state_dict), g["checksum"], rtol=1e-6), "synthetic scene drifted from the fixture"
    rays = T(g["rays"])
    is_train = bool(g["is_train"])
    u_c = T(g["u_coarse"]) if "u_coarse" in g else None
    u_f = T(g["u_fine"]) if "u_fine" in g else None
    cfg = oracle_cfg(scene, **okw)
    with torch.no_grad():
        out = O.render(scene.state_dict, cfg, rays, is_train, u_c, u_f, emission=scene.emission)
    e_rgb = np.abs(out[0].numpy() - g["rgb"]).max()
    e_dep = np.abs(out[1].numpy() - g["depth"]).max()
    assert out[4].shape == g["alpha"].shape
    print(f"{name}: oracle vs reference rgb {e_rgb:.2e} depth {e_dep:.2e}")
    assert e_rgb <= 1e-5
    assert e_dep <= 1e-4 * scene.near_far[1]
    if "bg" in g:
        assert np.abs(out[2].numpy() - g["bg"]).max() <= 1e-5
        assert np.abs(out[3].numpy() - g["env"]).max() <= 2e-6
    else:
        assert out[2] is None and out[3] is None


@pytest.mark.parametrize("name", FAST)
def test_render_tiny(name):
    _run_case(name)


@pytest.mark.parametrize("name", SLOW)
def test_render_baseline_shapes(name):
    _run_case(name)


@pytest.mark.parametrize("name", ["render_tiny_train_grad", "render_tiny_env_train_grad", "render_tiny_plain_train_grad"])
def test_gradients(name):
    """autograd through the oracle against the reference's own .grad (all 38/39 parameter tensors)."""
    skw, okw = RENDER_CASES[name]
    g = load_golden(name)
    scene = scene_for(skw)
    sd = {k: v.clone().requires_grad_(True) for k, v in scene.state_dict.items()}
    em = scene.emission.clone().requires_grad_(True) if scene.emission is not None else None
    out = O.render(sd, oracle_cfg(scene, **okw), T(g["rays"]), True, T(g["u_coarse"]), T(g["u_fine"]), emission=em)
    loss = (out[0] * T(g["w_rgb"])).sum() + (out[4] * T(g["w_alpha"])).sum()
    if em is not None:
        loss = loss + (out[2] * T(g["w_bg"])).sum() + (out[3] * T(g["w_env"])).sum()
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) <= 1e-4 * max(1., abs(float(g["loss"])))
    for k, v in sd.items():
        ref = g["grad:" + k]
        got = v.grad.numpy() if v.grad is not None else np.zeros_like(ref)
        scale = max(np.abs(ref).max(), 1e-6)
        assert np.abs(got - ref).max() <= 2e-4 * scale, k
    if em is not None:
        ref = g["grad:envmap.emission"]
        assert np.abs(em.grad.numpy() - ref).max() <= 2e-4 * max(np.abs(ref).max(), 1e-6)


def test_upsampled_factors_and_render_after_upsampling():
    """the oracle's restatement of the coarse-to-fine step (train.py:371-377) reproduces the reference's
    24 resampled factor tensors to 1 ulp and, with the r0 = 0.05 reset of set_resolution, its render on the new grid."""
    import dataclasses
    g = load_golden("upsample_tiny")
    scene = scene_for(dict(n_voxels=int(g["n_voxels"]), seed=int(g["seed"])))
    assert np.allclose(checksum(scene.state_dict), g["checksum"], rtol=1e-6)
    assert list(g["grid_old"]) == scene.grid
    reso = [int(v) for v in g["grid_new"]]
    assert reso == O.yinyang_resolution(28 ** 3)
    new = O.upsample_factors(scene.state_dict, reso, O.max_corner_radius(scene.aabb), scene.r0, scene.grid[0])
    n = 0
    for k in g.files:
        if k.startswith("sd:"):
            ref = g[k]
            assert tuple(new[k[3:]].shape) == ref.shape
            assert np.abs(new[k[3:]].numpy() - ref).max() <= 1e-6, k         # values up to ~3: <= 2 ulp
            n += 1
    assert n == 24
    up_scene = dataclasses.replace(scene, grid=reso, r0=float(g["r0_after"]), state_dict=new)
    assert up_scene.r0 == 0.05
    with torch.no_grad():
        out = O.render(new, oracle_cfg(up_scene), T(g["rays"]), False)
    assert np.abs(out[0].numpy() - g["rgb"]).max() <= 1e-5
    assert np.abs(out[1].numpy() - g["depth"]).max() <= 1e-4 * scene.near_far[1]


def test_occupancy_mask_family():
    """getDenseAlpha / updateAlphaMask / compute_alpha (EgoNeRF.py:438-489, tensorBase.py:421-436) restated
    by the oracle reproduce the reference's lattice alphas, binary volumes and masked alphas."""
    g = load_golden("alpha_mask_tiny")
    scene = scene_for(TINY)
    assert np.allclose(checksum(scene.state_dict), g["checksum"], rtol=1e-6)
    grid = [int(v) for v in g["grid"]]
    a_yin, a_yang = O.dense_alpha(scene.state_dict, grid, T(g["step"]), scene.density_shift)
    assert np.abs(a_yin.numpy() - g["alpha_yin"]).max() <= 1e-6
    assert np.abs(a_yang.numpy() - g["alpha_yang"]).max() <= 1e-6
    vols = O.alpha_mask_volumes(a_yin, a_yang, float(g["thres"]))
    assert np.array_equal(vols[0].numpy(), g["mask_yin"]) and np.array_equal(vols[1].numpy(), g["mask_yang"])
    assert 0.3 < vols[0].mean() < 0.7                                   # a mask that actually rejects something
    ma = O.compute_alpha(scene.state_dict, T(g["coords7"]), T(g["step"]), scene.density_shift, mask=vols)
    assert np.array_equal(ma.numpy() == 0, g["masked_alpha"] == 0)
    assert np.abs(ma.numpy() - g["masked_alpha"]).max() <= 1e-6


def test_plain_ladder_coordinates():
    """Without interval_th: normalize_coord = closed form on r0 * ratio^k (coordinates.py:132-156), halved ladder under
    `downsample=2`; includes r < r0, r = r0, the origin and radii beyond the grid."""
    g = load_golden("kat_coords_plain")
    aabb, grid, r0 = T(g["aabb"]), [int(v) for v in g["grid"]], float(g["r0"])
    far_r = O.max_corner_radius(aabb)
    assert float(far_r) == float(g["far_r"])
    r, a, b, is_yang, _ = O.cart_to_yinyang(T(g["points"]), O.scene_center(aabb))
    an, bn = O.normalize_angles(a, b)
    for key, ds in (("normalized", None), ("normalized_coarse", 2)):
        rn = O.normalize_radius_plain(r, far_r, r0, grid[0], downsample=ds)
        ref = T(g[key])
        act = torch.where(is_yang[:, None], ref[:, 3:6], ref[:, 0:3])
        assert torch.equal(ref[:, 6] != 0, is_yang)
        ok = torch.isfinite(act[:, 0])                      # the origin: log(0) -> the reference itself yields r/r0 = 0
        assert ok.all()
        assert (torch.stack([rn, an, bn], -1) - act).abs().max() <= 2e-6, key


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree only exists in the build container")
def test_committed_fixtures_are_what_the_reference_produces(tmp_path):
    """Re-runs oracle/make_golden.py (the UNMODIFIED reference, imported from /root/reference) into a scratch directory and
    compares every array of every committed fixture bit for bit: the goldens are reproducible reference outputs, not
    hand-edited numbers.  Never runs on the GPU box (no reference tree there)."""
    import subprocess
    import sys
    code = ("import sys, warnings; warnings.simplefilter('ignore'); sys.path.insert(0, %r);"
            "import oracle.make_golden as G; G.OUT = %r; G.main()") % (ROOT, str(tmp_path))
    run = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert run.returncode == 0, run.stderr[-2000:]
    committed = sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz"))
    assert committed == sorted(f for f in os.listdir(tmp_path) if f.endswith(".npz"))
    for f in committed:
        a, b = np.load(os.path.join(GOLDEN, f)), np.load(os.path.join(str(tmp_path), f))
        assert set(a.files) == set(b.files), f
        for k in a.files:
            assert a[k].shape == b[k].shape and np.array_equal(a[k], b[k], equal_nan=a[k].dtype.kind == "f"), (f, k)


def test_regulariser_mirror_matches_the_reference_fixture():
    """regularisers_tiny.npz (values + gradients from the unmodified reference, utils.py:155-183, EgoNeRF.py:189-229): the
    plain-torch regulariser methods of the drop-in module -- what an unchanged train.py:288-305 calls -- give the same values
    and, through autograd, the same gradients on the same state dict."""
    from egonerf_b200.scene_io import model_from_scene
    from tests.helpers import TINY, scene_for, checksum
    g = load_golden("regularisers_tiny")
    scene = scene_for(TINY)
    assert np.allclose(checksum(scene.state_dict), g["checksum"], rtol=1e-6)
    model = model_from_scene(scene, "cpu")
    tv = lambda x: 2 * (torch.pow(x[:, :, 1:, :] - x[:, :, :-1, :], 2).sum() / x[:, :, 1:, :].numel()
                        + torch.pow(x[:, :, :, 1:] - x[:, :, :, :-1], 2).sum() / x[:, :, :, 1:].numel()) / x.shape[0]
    vals = [model.TV_loss_density(tv), model.TV_loss_app(tv), model.density_L1(), model.vector_comp_diffs()]
    assert np.abs(torch.stack(vals).detach().numpy() - g["values"]).max() <= 1e-6 * np.abs(g["values"]).max() + 1e-9
    w = g["weights"]
    (float(w[0]) * vals[0] + float(w[1]) * vals[1] + float(w[2]) * vals[2]).backward()
    for name, p in model.named_parameters():
        if "plane" in name or "line" in name:
            ref = g["grad:" + name]
            got = np.zeros_like(ref) if p.grad is None else p.grad.numpy()
            assert np.abs(got - ref).max() <= 1e-6 * max(np.abs(ref).max(), 1e-12) + 1e-12, name
