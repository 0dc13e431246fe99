"""Quick GPU check while iterating on the fused fine pass: stage times at cfg2 (65 536 rays, 300^3 grid) + rgb / alpha / depth
of the throughput mode against the library's own fp32-equivalent mode on the same rays (a correctness tripwire, not the
parity test -- that is tests/test_gpu_tc.py).  Usage: python scripts/fused_quick.py [tag]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from egonerf_b200.scene_io import RENDER_KW, model_from_scene          # noqa: E402
from egonerf_b200.synthetic import make_rays, make_scene               # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "run"
dev = torch.device("cuda:0")
scene = make_scene(n_voxels=float(os.environ.get("EGN_QUICK_VOXELS", 27e6)))
model = model_from_scene(scene, dev)
rays = make_rays(65536, 'isotropic', seed=1000).to(dev)
out = {"tag": tag}
with torch.no_grad():
    model.mlp_mode, model.table_dtype = "tc_split", "f32"
    ref = model(rays[:8192], is_train=False, **RENDER_KW)
    z = model.sample_depths(rays[:8192], is_train=False, n_coarse=128, n_fine=128)
    ref_z = model(rays[:8192], is_train=False, z_vals=z, **RENDER_KW)
    model.mlp_mode, model.table_dtype = "tc_f16", "bf16"
    got = model(rays[:8192], is_train=False, **RENDER_KW)
    got_z = model(rays[:8192], is_train=False, z_vals=z, **RENDER_KW)
    out["rgb_linf_vs_parity_mode"] = float((got[0] - ref[0]).abs().max())
    out["alpha_linf"] = float((got[4] - ref[4]).abs().max())
    out["depth_rel"] = float(((got[1] - ref[1]).abs() / ref[1].abs().clamp_min(1e-3)).max())
    out["rgb_linf_fixed_depths"] = float((got_z[0] - ref_z[0]).abs().max())
    for _ in range(2):
        model.stage_times(rays, repeats=3, **RENDER_KW)
    st = model.stage_times(rays, repeats=10, **RENDER_KW)
    out["stage_ms"] = [round(x, 4) for x in st]
    # training-shape forward (features saved, separate compositing kernel) + backward
    tr = rays[:16384]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
with torch.enable_grad():
    for i in range(6):
        if i == 2:
            ev[0].record()
        for p in model.parameters():
            p.grad = None
        rgb = model(tr, is_train=True, seed=7, **RENDER_KW)[0]
        (rgb ** 2).mean().backward()
    ev[1].record()
    torch.cuda.synchronize()
    out["train_fwd_bwd_ms_16384"] = round(ev[0].elapsed_time(ev[1]) / 4, 4)
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"fused_quick_{tag}.json"), "w") as f:
    f.write(json.dumps(out) + "\n")
