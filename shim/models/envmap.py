"""Drop-in for models/envmap.py:6-37."""
from egonerf_b200.models.envmap import EnvironmentMap   # noqa: F401
