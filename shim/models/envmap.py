"""Drop-in for models/envmap.py:6-37 (the reference's other models import `EnvironmentMap` from here too)."""
import _egn_locate

_egn_locate.ensure_package_importable()
from egonerf_b200.models.envmap import EnvironmentMap   # noqa: E402,F401
