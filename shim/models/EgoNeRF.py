"""Drop-in for models/EgoNeRF.py:27 -- same class name, constructor kwargs, parameter names and forward signature.
`YinYangAlphaGridMask` (EgoNeRF.py:11-24) is mirrored too; the names the reference file pulls in with
`from models.tensorBase import *` stay importable from here."""
import _egn_locate

_egn_locate.ensure_package_importable()
if _egn_locate.reference_root(required=False) is not None:
    from models.tensorBase import *                                          # noqa: E402,F401,F403  (reference file)
from egonerf_b200.models.EgoNeRF import EgoNeRF, YinYangAlphaGridMask       # noqa: E402,F401
