"""Debug helper (GPU box): per-stage comparison of libegn_b200's workspace against the oracle's intermediates."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from egonerf_b200 import _lib
from egonerf_b200.scene_io import model_from_scene, RENDER_KW
from egonerf_b200.models import EgoNeRF as EM
from tests.helpers import RENDER_CASES, T, load_golden, oracle_cfg, scene_for, stable_rays
from oracle import egn_oracle as O

name = sys.argv[1] if len(sys.argv) > 1 else "render_128_eval"
skw, okw = RENDER_CASES[name]
g = load_golden(name); scene = scene_for(skw)
rays = T(g["rays"]); is_train = bool(g["is_train"])
u_c = T(g["u_coarse"]) if "u_coarse" in g else None
u_f = T(g["u_fine"]) if "u_fine" in g else None
model = model_from_scene(scene)
# capture the workspace
orig = torch.empty
keep = {}
def spy(*a, **k):
    t = orig(*a, **k)
    if k.get("dtype") == torch.uint8: keep["ws"] = t
    return t
torch.empty = spy
kw = dict(RENDER_KW); kw.update(okw)
with torch.no_grad():
    out = model(rays.cuda(), is_train=is_train, u_coarse=None if u_c is None else u_c.cuda(), u_fine=None if u_f is None else u_f.cuda(), **kw)
torch.empty = orig
torch.cuda.synchronize()
cfg = oracle_cfg(scene, **okw)
with torch.no_grad():
    ref, aux = O.render(scene.state_dict, cfg, rays, is_train, u_c, u_f, emission=scene.emission, want_aux=True)
N = rays.shape[0]; S = aux["z"].shape[1]; M = N * S
ws = keep["ws"].cpu().numpy()
al = lambda b: (b + 255) // 256 * 256
off = 0
def take(nfl):
    global off
    a = np.frombuffer(ws[off:off + nfl * 4].tobytes(), dtype=np.float32); off += al(nfl * 4); return a
z = take(M).reshape(N, S); fs = take(M).reshape(N, S); rgbs = take(M * 3).reshape(N, S, 3); wgt = take(M).reshape(N, S); take(N); take(N * 3); feat = take(M * 28).reshape(N, S, 28)
ez = np.abs(z - aux["z"].numpy())
print("z err max", ez.max(), "rays with z err>1e-4:", (ez.max(1) > 1e-4).sum())
sig = torch.nn.functional.softplus(torch.from_numpy(fs) + scene.density_shift).numpy()
es = np.abs(sig - aux["sigma"].numpy()); print("sigma err max", es.max(), np.unravel_index(es.argmax(), es.shape))
ef = np.abs(feat[..., :scene.app_dim] - aux["feat"].numpy()); print("feat err max", ef.max())
er = np.abs(rgbs - aux["rgb_samples"].numpy()); print("rgb_s err max", er.max())
ew = np.abs(wgt - aux["weight"].numpy()); print("weight err max", ew.max())
e = np.abs(out[0].cpu().numpy() - ref[0].numpy()).max(1)
eg = np.abs(out[0].cpu().numpy() - g["rgb"]).max(1)
print("rgb err vs oracle max", e.max(), "vs golden", eg.max(), "ray", e.argmax())
bad = np.argsort(-e)[:5]
for b in bad:
    print("ray", b, "rgb err", e[b], "z err", ez[b].max(), "at", ez[b].argmax(), "sigma err", es[b].max(), "at", es[b].argmax(), "feat err", ef[b].max(), "rgbs err", er[b].max(), "w err", ew[b].max(), "margin", float(aux["margin"][b]))
    j = es[b].argmax()
    print("   sample", j, "z", z[b, j], aux["z"][b, j].item(), "sigma", sig[b, j], aux["sigma"][b, j].item(), "yang", bool(aux["is_yang"][b, j]), "coords", aux["coords"][b, j].numpy())
