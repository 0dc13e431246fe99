"""CPU: the shim/ directory exposes the reference's module paths and names (train.py:6,11-12 imports), and the drop-in
EgoNeRF keeps the reference's parameter names / shapes (checkpoint + optimiser-group compatibility)."""
import importlib
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_module_paths_and_names():
    """Without a reference checkout behind it the shim still serves the names of the B200 path, and says why anything
    else is missing (the GPU box has no /root/reference)."""
    code = ("import sys; sys.path.insert(0, %r);"
            "import renderer, models; from models.EgoNeRF import EgoNeRF; from models.envmap import EnvironmentMap;"
            "from models import coordinates_dict;"
            "assert callable(renderer.volume_renderer) and renderer.OctreeRender_trilinear_fast is renderer.volume_renderer;"
            "assert 'yinyang' in coordinates_dict; print('ok', EgoNeRF.__module__)\n"
            "try:\n    renderer.evaluation\nexcept ImportError as e:\n    print('missing:', e)") % (os.path.join(ROOT, "shim"),)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp",
                         env={k: v for k, v in os.environ.items() if k != "EGONERF_REFERENCE"})
    assert out.returncode == 0, out.stderr
    assert "ok egonerf_b200.models.EgoNeRF" in out.stdout
    assert "missing:" in out.stdout and "reference checkout" in out.stdout


REF = os.environ.get("EGONERF_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "train.py")), reason="needs the reference checkout")
def test_unmodified_train_py_runs_through_the_shim(tmp_path):
    """SURVEY.md 8(b) / INTEGRATION.md A: with shim/ ahead of the reference checkout, the reference's OWN train.py imports
    (train.py:1-20 executed verbatim), builds args from its own config chain through its own opt.py, and its unmodified
    `train()` constructs dataset -> coordinates -> EgoNeRF -> Adam (train.py:118-186) on the drop-in classes and reaches the
    first `renderer(...)` call (train.py:253), which on a machine without a GPU must stop with the drop-in's own
    no-CPU-fallback error (on a GPU it trains; tests/test_gpu_dropin.py restates that loop for the box without the reference)."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_harness.py"), REF, str(tmp_path)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = {l.split()[0]: l for l in out.stdout.splitlines() if l and l.split()[0].isupper()}
    # the import block resolved: path functions -> egonerf_b200, everything else -> the reference's own modules
    assert lines["IMPORT_BLOCK_OK"].split()[1:] == ["egonerf_b200.renderer", "_egn_reference_renderer", "egonerf_b200.models.EgoNeRF",
                                                    "models.tensoRF", "egonerf_b200.models.coordinates"]
    # args as opt.py builds them from barbershop/default.txt -> common.txt -> common_indoor.txt -> EgoNeRF/common.txt
    assert lines["ARGS_OK"].split()[1:3] == ["EgoNeRF", "yinyang"]
    assert "[16, 16, 16] [48, 48, 48] [0.01, 15.0] 0.03 -8.0 True True 128 128 MLP_Fea 2" in lines["ARGS_OK"]
    assert lines["MODEL_OK"].split()[1] == "egonerf_b200.models.EgoNeRF" and lines["MODEL_OK"].split()[-1] == "11"
    import torch
    if torch.cuda.is_available():
        assert "TRAIN_DONE" in lines and "dropin.th" in lines["TRAIN_DONE"]
    else:
        assert "no CPU fallback" in lines["TRAIN_STOPPED_AT"]


def test_parameter_names_shapes_and_optimizer_groups_match_the_reference():
    from egonerf_b200.scene_io import model_from_scene
    from egonerf_b200.synthetic import make_scene
    import inspect
    from egonerf_b200.renderer import volume_renderer
    scene = make_scene(n_voxels=40 ** 3, seed=8, envmap_h=16, near_far=(0.1, 300.), r0=0.05, density_shift=-10.)
    model = model_from_scene(scene, "cpu")
    sd = model.state_dict()
    assert list(sd.keys()) != [] and set(sd.keys()) == set(scene.state_dict.keys())      # reference key set (EgoNeRF.py:96-122)
    g = model.gridSize.tolist()
    assert tuple(sd["density_plane_yin.0"].shape) == (1, 16, g[1], g[0])                  # (1, C, G[m1], G[m0])
    assert tuple(sd["app_line_yang.2"].shape) == (1, 48, g[0], 1)                          # (1, C, G[v], 1)
    assert tuple(sd["basis_mat_yin.weight"].shape) == (27, 144)
    groups = model.get_optparam_groups(0.02, 0.001, 0.1)                                  # EgoNeRF.py:139-156
    assert len(groups) == 5 * 2 + 1 + 1
    assert [gr["lr"] for gr in groups] == [0.02] * 4 + [0.001] + [0.02] * 4 + [0.001] + [0.001, 0.1]
    torch.optim.Adam(groups, betas=(0.9, 0.99))
    # same keyword surface as renderer.py:11-15
    ref_args = ["rays", "model", "chunk", "n_coarse", "n_fine", "ndc_ray", "white_bg", "is_train", "exp_sampling", "device",
                "empty_gpu_cache", "pretrain_envmap", "pivotal_sample_th", "resampling", "use_coarse_sample", "interval_th"]
    assert list(inspect.signature(volume_renderer).parameters) == ref_args
    kw = model.get_kwargs()
    for key in ("aabb", "gridSize", "density_n_comp", "appearance_n_comp", "app_dim", "density_shift", "distance_scale",
                "near_far", "shadingMode", "view_pe", "fea_pe", "featureC", "coordinates", "use_envmap", "envmap"):
        assert key in kw                                                                    # tensorBase.py:241-268


def test_surface_train_py_touches_exists_with_reference_signatures():
    """Every model / coordinates attribute the reference's train.py and renderer.py touch (train.py:96-97,123,189,253-270,
    356-380; renderer.py:24-31) exists on the drop-in objects, with the reference's argument names."""
    import inspect
    from egonerf_b200.models.EgoNeRF import EgoNeRF, YinYangAlphaGridMask
    from egonerf_b200.models.coordinates import YinYangSphericalCoords, coordinates_dict
    for name in ("forward", "get_optparam_groups", "save", "load", "get_kwargs", "update_coarse_sigma_grid", "upsample_volume_grid",
                 "up_sampling_VM", "updateAlphaMask", "getDenseAlpha", "compute_alpha", "vector_comp_diffs", "density_L1",
                 "TV_loss_density", "TV_loss_app", "feature2density", "compute_densityfeature", "compute_coarse_densityfeature",
                 "compute_appfeature", "update_stepSize", "init_svd_volume", "init_one_svd"):
        assert callable(getattr(EgoNeRF, name)), name
    fwd = list(inspect.signature(EgoNeRF.forward).parameters)
    assert fwd[1:13] == ["rays_chunk", "white_bg", "is_train", "ndc_ray", "n_coarse", "n_fine", "exp_sampling", "pretrain_envmap",
                         "pivotal_sample_th", "resampling", "use_coarse_sample", "interval_th"]            # EgoNeRF.py:491-495
    assert list(inspect.signature(EgoNeRF.compute_alpha).parameters)[1:] == ["norm_locs", "length"]        # tensorBase.py:421
    assert list(inspect.signature(EgoNeRF.up_sampling_VM).parameters)[1:] == ["plane_coef", "line_coef", "res_target"]
    assert list(inspect.signature(YinYangAlphaGridMask.__init__).parameters)[1:] == ["device", "alpha_volume_yin", "alpha_volume_yang"]
    for name in ("N_to_reso", "set_resolution", "update_aabb", "up_sampling_VM", "normalize_r"):
        assert callable(getattr(YinYangSphericalCoords, name)), name
    assert list(inspect.signature(YinYangSphericalCoords.__init__).parameters)[1:] == ["device", "aabb", "exp_r", "N_voxel", "r0",
                                                                                      "interval_th"]     # train.py:122-124
    assert list(inspect.signature(YinYangSphericalCoords.up_sampling_VM).parameters)[1:] == ["weights", "res_target", "ids"]
    # both ladder flavours construct on the host (no GPU needed for the ladders)
    aabb = torch.tensor([[-15.5] * 3, [15.5] * 3])
    for ith in (True, False):
        co = coordinates_dict["yinyang"]("cpu", aabb, exp_r=True, N_voxel=40 ** 3, r0=0.03, interval_th=ith)
        assert co.r_knots().shape[0] == co.N_r + (1 if ith else 3)
        co.set_resolution(co.N_to_reso(64 ** 3, aabb))
        assert co.r0 == 0.05                                       # the reference's set_resolution default (coordinates.py:214)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree only exists in the build container")
def test_reference_checkpoint_kwargs_unpickle_through_the_shim(tmp_path):
    """A reference checkpoint pickles its coordinates and envmap OBJECTS into `kwargs` (tensorBase.py:241-268); loaded with
    shim/ on the path they must come back as the drop-in classes, carrying the reference's scalars and ladders."""
    path = str(tmp_path / "kwargs.th")
    make = ("import sys, types, torch; sys.dont_write_bytecode = True; sys.path.insert(0, %r);"
            "from oracle import ref_harness; cd, *_ = ref_harness.import_reference();"
            "from models.envmap import EnvironmentMap;"
            "aabb = torch.tensor([[-15.5] * 3, [15.5] * 3]);"
            "co = cd['yinyang']('cpu', aabb, exp_r=True, N_voxel=40 ** 3, r0=0.03, interval_th=True);"
            "assert type(co).__module__ == 'models.coordinates';"
            "torch.save({'kwargs': {'coordinates': co, 'envmap': EnvironmentMap(h=4, init_strategy='zero', device='cpu')}}, %r)"
            ) % (ROOT, path)
    run = subprocess.run([sys.executable, "-c", make], capture_output=True, text=True)
    assert run.returncode == 0, run.stderr[-1500:]
    load = ("import sys, torch; sys.path.insert(0, %r); sys.path.insert(0, %r);"
            "kw = torch.load(%r, map_location='cpu', weights_only=False)['kwargs'];"
            "from egonerf_b200.models.coordinates import YinYangSphericalCoords; from egonerf_b200.models.envmap import EnvironmentMap;"
            "from oracle import egn_oracle as O;"
            "co = kw['coordinates']; assert type(co) is YinYangSphericalCoords, type(co);"
            "assert type(kw['envmap']) is EnvironmentMap and tuple(kw['envmap'].emission.shape) == (3, 8, 4);"
            "assert [co.N_r, co.N_theta, co.N_phi] == [20, 22, 64] and co.r0 == 0.03 and co.interval_th;"
            "aabb = torch.tensor([[-15.5] * 3, [15.5] * 3]);"
            "assert torch.equal(co.r_knots(), O.r_reference_grid(O.max_corner_radius(aabb), 0.03, 20)); print('ok')"
            ) % (ROOT, os.path.join(ROOT, "shim"), path)
    run = subprocess.run([sys.executable, "-c", load], capture_output=True, text=True)
    assert run.returncode == 0 and "ok" in run.stdout, run.stderr[-1500:]


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree only exists in the build container")
def test_checkpoint_saved_by_the_reference_loads_like_train_py_does(tmp_path):
    """`model.save` of the UNMODIFIED reference (EgoNeRF.py:158-172), then the load sequence of train.py:155-160 with shim/
    on the path: torch.load -> kwargs -> EgoNeRF(**kwargs) -> model.load(ckpt).  Parameters, envmap and scalars arrive."""
    path = str(tmp_path / "ref.th")
    save = ("import sys, torch; sys.dont_write_bytecode = True; sys.path.insert(0, %r);"
            "from oracle import ref_harness; ref_harness.import_reference();"
            "from oracle.make_golden import build_reference; from egonerf_b200.synthetic import make_scene;"
            "scene = make_scene(n_voxels=40 ** 3, seed=8, envmap_h=16, near_far=(0.1, 300.), r0=0.05, density_shift=-10.);"
            "co, model = build_reference(scene); assert type(model).__module__ == 'models.EgoNeRF'; model.save(%r, global_step=123)"
            ) % (ROOT, path)
    run = subprocess.run([sys.executable, "-c", save], capture_output=True, text=True)
    assert run.returncode == 0, run.stderr[-1500:]
    load = ("import sys, torch; sys.path.insert(0, %r); sys.path.insert(0, %r);"
            "from models.EgoNeRF import EgoNeRF;"
            "ckpt = torch.load(%r, map_location='cpu', weights_only=False);"
            "kwargs = ckpt['kwargs']; kwargs.update({'device': 'cpu'});"
            "model = EgoNeRF(**kwargs); assert model.load(ckpt) == 123 and type(model).__module__ == 'egonerf_b200.models.EgoNeRF';"
            "from egonerf_b200.synthetic import make_scene;"
            "scene = make_scene(n_voxels=40 ** 3, seed=8, envmap_h=16, near_far=(0.1, 300.), r0=0.05, density_shift=-10.);"
            "sd = model.state_dict(); assert all(torch.equal(sd[k], v) for k, v in scene.state_dict.items());"
            "assert torch.equal(model.envmap.emission.detach(), scene.emission);"
            "assert model.near_far == [0.1, 300.0] and model.density_shift == -10.0 and model.coordinates.r0 == 0.05;"
            "assert model.gridSize.tolist() == scene.grid; print('ok')") % (ROOT, os.path.join(ROOT, "shim"), path)
    run = subprocess.run([sys.executable, "-c", load], capture_output=True, text=True)
    assert run.returncode == 0 and "ok" in run.stdout, run.stderr[-1500:]


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree only exists in the build container")
@pytest.mark.parametrize("interval_th", [True, False])
def test_upsampling_positions_match_the_reference_ladders(interval_th):
    """Host half of the coarse-to-fine step: the source positions handed to `egn_resample_factor` for the r axis equal the
    grid_sample coordinates the reference builds in GenericSphericalCoords.up_sampling_VM (coordinates.py:238-250), with and
    without interval_th; angular axes follow F.interpolate's align_corners rule."""
    code = ("import sys, torch; sys.dont_write_bytecode = True; sys.path.insert(0, %r);"
            "from oracle import ref_harness; cd, *_ = ref_harness.import_reference();"
            "from extra.test_exp_r import index2r;"
            "aabb = torch.tensor([[-15.5] * 3, [15.5] * 3]); ith = %r;"
            "co = cd['yinyang']('cpu', aabb, exp_r=True, N_voxel=40 ** 3, r0=0.03, interval_th=ith);"
            "n = 34; ratio = pow(co.far[0] / co.r0, 1 / (n - 1));"
            "g = index2r(co.r0, ratio, torch.arange(n));"
            "iv = g[1:] - g[:-1]; cum = torch.cumsum(iv, 0); k = int((iv <= co.r0).sum());"
            "gi = g.clone(); gi[:k + 1] = torch.arange(k + 1) * co.r0; gi[k + 1:] = g[k + 1:] + co.r0 * k - cum[k - 1];"
            "un = torch.zeros(n); un[1:] = co.r0 * torch.pow(ratio, torch.arange(n - 1));"
            "rs = co.normalize_r(gi if ith else un) * 2 - 1;"
            "print(' '.join(repr(float(v)) for v in ((rs + 1) / 2) * (co.N_r - 1)))") % (ROOT, interval_th)
    run = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert run.returncode == 0, run.stderr[-1500:]
    ref = torch.tensor([float(v) for v in run.stdout.split()])
    from egonerf_b200.models.coordinates import YinYangSphericalCoords
    aabb = torch.tensor([[-15.5] * 3, [15.5] * 3])
    co = YinYangSphericalCoords("cpu", aabb, exp_r=True, N_voxel=40 ** 3, r0=0.03, interval_th=interval_th)
    mine = co.up_sampling_positions(0, co.N_r, 34)
    assert mine.shape == ref.shape and (mine - ref).abs().max() <= 2e-5, (mine - ref).abs().max()
    ang = co.up_sampling_positions(1, co.N_theta, 37)
    assert ang[0] == 0 and abs(float(ang[-1]) - (co.N_theta - 1)) < 1e-5 and torch.all(ang[1:] > ang[:-1])


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree only exists in the build container")
def test_regularisers_equal_the_reference_values():
    """vector_comp_diffs / density_L1 / TV_loss_density / TV_loss_app (EgoNeRF.py:189-230 with TVLoss of
    utils.py:155-171) are plain torch on the Parameters — same numbers as the unmodified reference on the same state dict."""
    code = ("import sys, torch; sys.dont_write_bytecode = True; sys.path.insert(0, %r);"
            "from oracle import ref_harness; ref_harness.import_reference();"
            "from oracle.make_golden import build_reference; from egonerf_b200.synthetic import make_scene;"
            "from utils import TVLoss;"
            "co, m = build_reference(make_scene(n_voxels=40 ** 3, seed=7)); tv = TVLoss();"
            "print(repr(float(m.vector_comp_diffs())), repr(float(m.density_L1())), repr(float(m.TV_loss_density(tv))),"
            " repr(float(m.TV_loss_app(tv))))") % ROOT
    run = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert run.returncode == 0, run.stderr[-1500:]
    ref = [float(v) for v in run.stdout.strip().splitlines()[-1].split()]
    from egonerf_b200.scene_io import model_from_scene
    from egonerf_b200.synthetic import make_scene

    class TV(torch.nn.Module):                                  # TVLoss, utils.py:155-171 (restated)
        def forward(self, x):
            n_h = x[:, :, 1:, :].numel() // x.shape[0]
            n_w = x[:, :, :, 1:].numel() // x.shape[0]
            dh = ((x[:, :, 1:, :] - x[:, :, :-1, :]) ** 2).sum()
            dw = ((x[:, :, :, 1:] - x[:, :, :, :-1]) ** 2).sum()
            return 2 * (dh / n_h + dw / n_w) / x.shape[0]

    m = model_from_scene(make_scene(n_voxels=40 ** 3, seed=7), "cpu")
    mine = [float(m.vector_comp_diffs()), float(m.density_L1()), float(m.TV_loss_density(TV())), float(m.TV_loss_app(TV()))]
    for a, b in zip(mine, ref):
        assert abs(a - b) <= 1e-6 * max(1.0, abs(b)), (mine, ref)
