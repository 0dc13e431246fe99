"""egonerf_b200 — B200 (sm_100a) volume-rendering path of EgoNeRF behind the reference's operator surface.

    from egonerf_b200.models.EgoNeRF import EgoNeRF
    from egonerf_b200.models import coordinates_dict
    from egonerf_b200.renderer import volume_renderer

All arithmetic of the path runs in the hand-written CUDA library `libegn_b200.so` (C ABI: include/egn.h).
There is no CPU fallback.
"""
__version__ = "0.1.0"
