"""Drop-in for the reference's `renderer` module: put this directory ahead of the reference checkout on PYTHONPATH
(INTEGRATION.md §A).  `volume_renderer` (renderer.py:11-79; alias `OctreeRender_trilinear_fast`) is the B200 path; every
other name -- `evaluation` (renderer.py:82-198), `evaluation_path` (:200-255) and what `from utils import *` brings in --
is the reference's own, loaded from the reference tree, so `train.py:6` (`from renderer import volume_renderer,
evaluation`) works unchanged.  `evaluation` receives the renderer as an argument (train.py:225,339), i.e. it drives the
B200 path too."""
import _egn_locate

_egn_locate.ensure_package_importable()
if _egn_locate.reference_root(required=False) is not None:
    _reference = _egn_locate.load_reference_module("renderer.py", "_egn_reference_renderer")
    _egn_locate.reexport(_reference, globals())
else:
    __getattr__ = _egn_locate.missing_name_hook("renderer")

from egonerf_b200.renderer import volume_renderer, OctreeRender_trilinear_fast   # noqa: E402,F401  (override)
