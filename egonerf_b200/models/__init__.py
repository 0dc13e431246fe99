from .coordinates import YinYangSphericalCoords, coordinates_dict   # noqa: F401
