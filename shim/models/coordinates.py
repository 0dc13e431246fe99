# replaces models/coordinates.py for the EgoNeRF path; also the class a reference checkpoint's pickled
# `kwargs['coordinates']` (models.coordinates.YinYangSphericalCoords) resolves to when it is loaded through the shim
from egonerf_b200.models.coordinates import YinYangSphericalCoords, coordinates_dict   # noqa
