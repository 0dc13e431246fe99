"""CPU: the C-ABI library builds/loads, exports every symbol include/egn.h declares, validates its arguments without a
GPU, and its host helpers reproduce the oracle's ladders.  No compute entry point is executed here."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from egonerf_b200 import _lib
from oracle import egn_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.isfile(_lib.LIB_PATH):
        from egonerf_b200.build import build
        build()
    return _lib.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "egn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(egn_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = _declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/egn.h but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype in egonerf_b200/_lib.py"
    assert set(_lib.PROTOTYPES) <= set(names), "ctypes binds symbols the header does not declare"
    assert lib.egn_abi_version() == _lib.ABI_VERSION


def test_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def _cfg(**kw):
    cfg = _lib.EgnConfig()
    cfg.grid[:] = [20, 22, 64]
    cfg.c_sigma, cfg.c_app, cfg.app_dim, cfg.shading = 16, 48, 27, 0
    cfg.view_pe, cfg.fea_pe, cfg.feature_c = 2, 2, 128
    cfg.n_coarse, cfg.n_fine, cfg.use_coarse_sample, cfg.resampling = 128, 128, 1, 1
    cfg.exp_sampling = 1
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def test_sizes_and_validation(lib):
    cfg = _cfg()
    assert lib.egn_samples_per_ray(cfg) == 256
    assert lib.egn_samples_per_ray(_cfg(use_coarse_sample=0)) == 128
    assert lib.egn_samples_per_ray(_cfg(resampling=0)) == 128
    nfl = lib.egn_table_floats(cfg)
    G = [20, 22, 64]
    fine = 2 * 64 * (G[0] * G[1] + G[0] * G[2] + G[1] * G[2] + sum(G))
    coarse = 2 * 16 * (10 * 11 + 10 * 32 + 11 * 32 + 10 + 11 + 32)
    assert fine + coarse <= nfl <= fine + coarse + 24 * 64          # sections padded to 256 B
    assert lib.egn_workspace_bytes(cfg, 1000) > lib.egn_workspace_bytes_eval(cfg, 1000) > 0
    # errors are reported through the status code + egn_last_error, never by crashing
    assert lib.egn_table_floats(_cfg(c_sigma=8)) == -1
    assert b"n_lamb_sigma" in lib.egn_last_error()
    bad = _cfg(n_coarse=100)
    z = (C.c_float * 4)()
    assert lib.egn_sample_rays(bad, C.addressof(z), C.addressof(z), 1, 0, None, None, 0, 0, C.addressof(z), None) != 0
    assert b"n_coarse" in lib.egn_last_error()
    with pytest.raises(RuntimeError, match="libegn_b200"):
        _lib.check(1)


@pytest.mark.parametrize("near,far,r0,n", [(0.01, 15., 0.03, 128), (0.1, 300., 0.05, 128), (0.01, 15., 0.03, 256)])
def test_host_sample_schedule_matches_oracle(lib, near, far, r0, n):
    out = (C.c_float * n)()
    assert lib.egn_host_sample_schedule(near, far, r0, n, out) == 0
    ref = O.sample_schedule(near, far, r0, n).numpy()
    assert np.abs(np.array(out) - ref).max() <= 4e-6 * ref.max()      # powf vs torch pow: <= a few ulp


@pytest.mark.parametrize("half,r0,n_r", [(15.5, 0.03, 150), (15.5, 0.03, 64), (300.5, 0.05, 150)])
def test_host_r_knots_matches_oracle(lib, half, r0, n_r):
    aabb = torch.tensor([[-half] * 3, [half] * 3])
    far_r = O.max_corner_radius(aabb)
    out = (C.c_float * (n_r + 1))()
    assert lib.egn_host_r_knots(float(far_r), r0, n_r, out) == 0
    ref = O.r_reference_grid(far_r, r0, n_r).numpy()
    assert np.abs(np.array(out) - ref).max() <= 4e-6 * ref.max()


def test_host_plain_ladders_match_oracle(lib):
    """The ladders of a run without --interval_th (EgnConfig.plain_ladders): C helpers and Python mirror against the oracle."""
    from egonerf_b200.models.coordinates import YinYangSphericalCoords, plain_sample_schedule
    n = 128
    out, ratio, r0p = (C.c_float * n)(), C.c_float(), C.c_float()
    assert lib.egn_host_plain_sample_schedule(0.01, 15., n, out, C.byref(ratio), C.byref(r0p)) == 0
    ref = O.plain_sample_schedule(0.01, 15., n)[0].numpy()
    assert np.abs(np.array(out) - ref).max() <= 4e-6 * ref.max()
    r, ratio_py, r0_py = plain_sample_schedule(0.01, 15., n)
    assert torch.equal(r, O.plain_sample_schedule(0.01, 15., n)[0])                 # eval depths: bit-exact
    assert abs(ratio.value - ratio_py) < 1e-7 and abs(r0p.value - r0_py) < 1e-8
    aabb = torch.tensor([[-15.5] * 3, [15.5] * 3])
    co = YinYangSphericalCoords("cpu", aabb, exp_r=True, N_voxel=40 ** 3, r0=0.03, interval_th=False)
    far_r = O.max_corner_radius(aabb)
    for ds, n_r in ((None, co.N_r), (2, co.N_r // 2)):
        knots = co.r_knots(downsample=ds)
        assert knots.shape[0] == n_r + 3 and knots[0] == 0 and abs(float(knots[1]) - 0.03) < 1e-9
        ck = (C.c_float * (n_r + 3))()
        assert lib.egn_host_plain_r_knots(float(far_r), 0.03, n_r, ck) == 0
        assert np.abs(np.array(ck) - knots.numpy()).max() <= 4e-6 * float(knots.max())
        # searching the knot ladder == the reference's closed form (log / trunc / pow), away from the knots themselves
        r = torch.rand(4000, generator=torch.Generator().manual_seed(1)) * float(far_r) * 1.05
        hi = torch.clamp(torch.searchsorted(knots, r, side="right"), 1, n_r + 2)
        mine = ((hi - 1) + (r - knots[hi - 1]) / (knots[hi] - knots[hi - 1])) / n_r * 2 - 1
        assert (mine - O.normalize_radius_plain(r, far_r, 0.03, co.N_r, downsample=ds)).abs().max() <= 2e-6


def test_python_mirror_ladders_are_bit_exact():
    """The drop-in modules feed the kernels the ladders built by the host mirror: these must equal the oracle's."""
    from egonerf_b200.models.coordinates import YinYangSphericalCoords, sample_schedule
    for half, r0, nvox, nf in ((15.5, 0.03, 27e6, (0.01, 15.)), (300.5, 0.05, 27e6, (0.1, 300.)), (15.5, 0.03, 128 ** 3, (0.01, 15.))):
        aabb = torch.tensor([[-half] * 3, [half] * 3])
        co = YinYangSphericalCoords("cpu", aabb, exp_r=True, N_voxel=nvox, r0=r0, interval_th=True)
        grid = O.yinyang_resolution(nvox)
        assert [co.N_r, co.N_theta, co.N_phi] == grid
        assert torch.equal(co.r_knots(), O.r_reference_grid(O.max_corner_radius(aabb), r0, grid[0]))
        assert torch.equal(sample_schedule(nf[0], nf[1], r0, 128), O.sample_schedule(nf[0], nf[1], r0, 128))


def test_no_cpu_fallback():
    """The product path refuses CPU tensors instead of silently computing somewhere else."""
    from egonerf_b200.scene_io import model_from_scene
    from egonerf_b200.synthetic import make_scene, make_rays
    scene = make_scene(n_voxels=40 ** 3, seed=7)
    model = model_from_scene(scene, "cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(make_rays(4), n_coarse=128, n_fine=128, exp_sampling=True, resampling=True)
    import egonerf_b200
    src = open(os.path.join(os.path.dirname(egonerf_b200.__file__), "models", "EgoNeRF.py")).read()
    assert "oracle" not in src


def test_ctypes_structs_mirror_the_header_field_by_field():
    """A reordered or missing field would silently shift every later argument: the ctypes mirror of EgnConfig / EgnParams /
    EgnOutputs must list the header's fields in the header's order."""
    text = open(os.path.join(ROOT, "include", "egn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)

    def fields(struct):
        body = text[text.index("typedef struct %s {" % struct):text.index("} %s;" % struct)].split("{", 1)[1]
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            first, *rest = decl.split(",")
            names.append(re.findall(r"(\w+)\s*(?:\[[^\]]*\])*\s*$", first)[0])
            names += [re.findall(r"(\w+)", r)[0] for r in rest]
        return names

    assert fields("EgnConfig") == [f[0] for f in _lib.EgnConfig._fields_]
    assert fields("EgnParams") == [f[0] for f in _lib.EgnParams._fields_]
    assert fields("EgnOutputs") == [f[0] for f in _lib.EgnOutputs._fields_]


def test_product_code_never_touches_the_oracle_or_the_reference_tree():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use oracle/; nothing shipped may read
    /root/reference (it does not exist on the GPU box)."""
    import ast
    pkg = os.path.join(ROOT, "egonerf_b200")
    offenders = []
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                if "/root/reference" in text or re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M):
                    offenders.append(os.path.join(base, f))
    for f in os.listdir(os.path.join(ROOT, "shim")):
        pass
    assert not offenders, offenders
    # bench.py: the oracle is imported inside oracle_rays_per_s (cpu_baseline / --impl reference) and nowhere else
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef):
            uses = any(isinstance(n, ast.ImportFrom) and n.module and n.module.startswith("oracle") for n in ast.walk(node))
            assert (not uses) or node.name == "oracle_rays_per_s", node.name
    top = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom))]
    assert not any(getattr(n, "module", "") and str(n.module).startswith("oracle") for n in top)
    assert "/root/reference" not in open(os.path.join(ROOT, "bench.py")).read().replace("Nothing here reads /root/reference", "")


def test_header_is_plain_c99_and_links_from_c(lib, tmp_path):
    """The drop-in boundary is a C ABI: include/egn.h must compile as strict C99 (no C++-isms, no torch types) and a plain C
    program must be able to size buffers and build the host ladders through libegn_b200.so (no GPU calls here)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "probe.c"
    src.write_text(r'''
#include <stdio.h>
#include "egn.h"
int main(void) {
    float z[128], k[151];
    EgnConfig cfg = {0};
    if (egn_abi_version() != EGN_ABI_VERSION) return 2;
    if (egn_host_sample_schedule(0.01f, 15.f, 0.03f, 128, z)) { printf("err %s\n", egn_last_error()); return 1; }
    if (egn_host_r_knots(26.846788f, 0.03f, 150, k)) return 1;
    cfg.grid[0] = 150; cfg.grid[1] = 172; cfg.grid[2] = 516; cfg.c_sigma = 16; cfg.c_app = 48; cfg.app_dim = 27;
    cfg.view_pe = cfg.fea_pe = 2; cfg.feature_c = 128; cfg.n_coarse = cfg.n_fine = 128;
    cfg.use_coarse_sample = cfg.resampling = 1; cfg.mlp_mode = EGN_MLP_TC_BF16;
    printf("%d %lld %lld %.6f %.6f\n", egn_samples_per_ray(&cfg), (long long)egn_table_floats(&cfg),
           (long long)egn_workspace_bytes_eval(&cfg, 4096), z[127], k[150]);
    cfg.c_sigma = 8;
    if (egn_table_floats(&cfg) != -1) return 3;
    return 0;
}
''')
    exe = tmp_path / "probe"
    libdir = os.path.dirname(_lib.LIB_PATH)
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src),
                         "-o", str(exe), "-L", libdir, "-legn_b200", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0, run.stdout + run.stderr
    S, nfl, ws, z_last, k_last = run.stdout.split()
    assert int(S) == 256 and int(nfl) == lib.egn_table_floats(_cfg(grid=(C.c_int32 * 3)(150, 172, 516))) and int(ws) > 0
    assert abs(float(z_last) + 0.01 - 15.5509796) < 2e-5            # SURVEY 8c known-answer value of the sample schedule
    assert abs(float(k_last) - float(O.r_reference_grid(torch.tensor(26.846788), 0.03, 150)[150])) < 1e-4
