"""Quick GPU timing of one training step's forward + backward at cfg2 (16 384 rays, 300^3 grid, throughput mode) for A/B
runs of library variants (EGN_B200_LIB): python scripts/train_quick.py [tag]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from egonerf_b200.scene_io import RENDER_KW, model_from_scene          # noqa: E402
from egonerf_b200.synthetic import make_rays, make_scene               # noqa: E402
from egonerf_b200.optim import TableAdam                               # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "run"
dev = torch.device("cuda:0")
model = model_from_scene(make_scene(n_voxels=27e6), dev)
model.mlp_mode, model.table_dtype = "tc_f16", "bf16"
opt = TableAdam(model, 0.02, 0.001, 0.1)
rays = make_rays(16384, 'isotropic', seed=2000).to(dev)
target = torch.rand(16384, 3, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
steps = int(os.environ.get("EGN_QUICK_STEPS", 20))
for i in range(steps + 3):
    if i == 3:
        ev[0].record()
    for p in model.parameters():
        p.grad = None
    opt.zero_grad()
    rgb = model(rays, is_train=True, seed=1234, **RENDER_KW)[0]
    torch.mean((rgb - target) ** 2).backward()
    opt.step()
    model.update_coarse_sigma_grid()
ev[1].record()
torch.cuda.synchronize()
out = {"tag": tag, "train_step_ms_16384": round(ev[0].elapsed_time(ev[1]) / steps, 4), "loss": float(torch.mean((rgb - target) ** 2))}
print(json.dumps(out))
