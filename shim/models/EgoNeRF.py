"""Drop-in for models/EgoNeRF.py:27 — same class name, constructor kwargs, parameter names and forward signature."""
from egonerf_b200.models.EgoNeRF import EgoNeRF   # noqa: F401
