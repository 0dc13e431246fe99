"""Drop-in for models/coordinates.py: every class of the reference module (the `Coordinates` base that
models/tensorBase.py:9 imports, the other eight coordinate systems) re-exported from the reference tree, with
`YinYangSphericalCoords` (coordinates.py:432-520) replaced by the B200 mirror.  Also the class a reference checkpoint's
pickled `kwargs['coordinates']` (models.coordinates.YinYangSphericalCoords) resolves to when loaded through the shim."""
import _egn_locate

_egn_locate.ensure_package_importable()
from egonerf_b200.models.coordinates import YinYangSphericalCoords          # noqa: E402  (override)

coordinates_dict = {'yinyang': YinYangSphericalCoords}
if _egn_locate.reference_root(required=False) is not None:
    _reference = _egn_locate.load_reference_module("models/coordinates.py", "_egn_reference_models_coordinates")
    _egn_locate.reexport(_reference, globals())
    coordinates_dict = {                                                     # models/__init__.py:5-15
        'xyz': _reference.CartesianCoords,
        'sphere': _reference.SphericalCoords,
        'balanced_sphere': _reference.BalancedSphericalCoords,
        'directional_sphere': _reference.DirectionalSphericalCoords,
        'directional_balanced_sphere': _reference.DirectionalBalancedSphericalCoords,
        'cylinder': _reference.CylindricalCoords,
        'euler_sphere': _reference.EulerSphericalCoords,
        'yinyang': YinYangSphericalCoords,
        'generic_sphere': _reference.GenericSphericalCoords,
    }
else:
    __getattr__ = _egn_locate.missing_name_hook("models.coordinates")
