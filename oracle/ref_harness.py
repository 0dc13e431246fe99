"""Import the UNMODIFIED reference (changwoonchoi/EgoNeRF, mounted at /root/reference) on CPU.

TEST INFRASTRUCTURE ONLY.  This module exists so that `oracle/make_golden.py` can run the real
reference in the build container and freeze its outputs under `tests/golden/`.  /root/reference does
not exist on the GPU box, so nothing on the product path, in `-m gpu` tests, `smoke()` or `bench.py`
may import this file.

The reference hard-imports a few packages that are absent here and unused by the volume-rendering
path: matplotlib (extra/test_exp_r.py:1), kornia (dataLoader/ray_utils.py:4),
imageio (renderer.py:1), skimage / plyfile (utils.py:7-8).  They are replaced by empty stubs.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("EGONERF_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "EgoNeRF.py"))


def import_reference():
    """Returns (coordinates_dict, EgoNeRF, volume_renderer, sample_pdf, raw2alpha) from the reference."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    sys.dont_write_bytecode = True  # the mount is read-only
    for name in ("matplotlib", "matplotlib.pyplot", "kornia", "imageio", "plyfile",
                 "skimage", "skimage.measure", "skimage.metrics", "lpips", "configargparse"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    k = sys.modules["kornia"]
    if not hasattr(k, "create_meshgrid"):
        k.create_meshgrid = lambda *a, **kw: None
    p = sys.modules["plyfile"]
    for attr in ("PlyData", "PlyElement"):
        if not hasattr(p, attr):
            setattr(p, attr, object)
    sk = sys.modules["skimage.measure"]
    if not hasattr(sk, "marching_cubes"):
        sk.marching_cubes = lambda *a, **kw: None
    # our own drop-in tree must not shadow the reference's `models` / `renderer`
    for mod in [m for m in sys.modules if m == "models" or m.startswith("models.") or m in ("renderer", "utils")]:
        del sys.modules[mod]
    sys.path.insert(0, REF_ROOT)
    try:
        from models import coordinates_dict
        from models.EgoNeRF import EgoNeRF
        from models.tensorBase import raw2alpha
        from dataLoader.ray_utils import sample_pdf
        from renderer import volume_renderer
    finally:
        sys.path.remove(REF_ROOT)
    return coordinates_dict, EgoNeRF, volume_renderer, sample_pdf, raw2alpha
