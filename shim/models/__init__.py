"""Drop-in for the reference's `models` package (models/__init__.py:5-15): only the Yin-Yang system is on the path."""
from egonerf_b200.models.coordinates import coordinates_dict, YinYangSphericalCoords   # noqa: F401
