"""Drop-in for the reference's `renderer` module (renderer.py:11-79): put this directory ahead of the reference on
PYTHONPATH.  Everything else the reference imports from `renderer` (evaluation, metrics) is
out of scope of this path and keeps coming from the reference tree."""
from egonerf_b200.renderer import volume_renderer, OctreeRender_trilinear_fast   # noqa: F401
