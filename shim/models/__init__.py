"""Drop-in for the reference's `models` package (models/__init__.py:1-15).  `models.EgoNeRF`, `models.coordinates` and
`models.envmap` resolve to the modules of this directory (the B200 path); every other submodule (`models.tensoRF`,
`models.tensorBase`, `models.sh`) resolves to the reference's file through the extended package path, so
`train.py:11` (`from models.tensoRF import TensorVM, TensorCP, raw2alpha, TensorVMSplit, AlphaGridMask`) works unchanged.
`coordinates_dict` keeps the reference's nine entries with 'yinyang' replaced."""
import os

import _egn_locate

_egn_locate.ensure_package_importable()
if _egn_locate.reference_root(required=False) is not None:
    __path__.append(os.path.join(_egn_locate.reference_root(), "models"))  # after this directory: shim modules win

from .coordinates import *                                                   # noqa: E402,F401,F403
from .coordinates import coordinates_dict, YinYangSphericalCoords            # noqa: E402,F401
