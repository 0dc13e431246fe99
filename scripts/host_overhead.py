"""Host-side cost of one forward call (Python + ctypes + launches), measured without waiting for the GPU."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from egonerf_b200.scene_io import RENDER_KW, model_from_scene
from egonerf_b200.synthetic import make_rays, make_scene
from egonerf_b200.renderer import volume_renderer
dev = torch.device("cuda:0")
model = model_from_scene(make_scene(n_voxels=27e6), dev)
model.mlp_mode, model.table_dtype = "tc_f16", "bf16"
rays = make_rays(65536, 'isotropic', seed=1).to(dev)
rays_h = rays.cpu().pin_memory()
with torch.no_grad():
    for _ in range(3):
        model(rays, is_train=False, **RENDER_KW)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model(rays, is_train=False, **RENDER_KW)
        ts.append(time.perf_counter() - t0)
    print("model() host time per call: median %.1f us, min %.1f us" % (sorted(ts)[10] * 1e6, min(ts) * 1e6))
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(50):
        model(rays, is_train=False, **RENDER_KW)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
    import io, contextlib
    ts = []
    for _ in range(10):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            out = volume_renderer(rays_h, model, chunk=65536, is_train=False, device=dev, **RENDER_KW)
        r = out[0].cpu()
        ts.append(time.perf_counter() - t0)
    print("volume_renderer from host rays + rgb to host: median %.3f ms" % (sorted(ts)[5] * 1e3))
