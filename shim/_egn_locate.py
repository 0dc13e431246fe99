"""Locates the UNMODIFIED reference tree (changwoonchoi/EgoNeRF) behind this shim directory.

The shim replaces exactly three modules of the reference -- `renderer.volume_renderer`, `models.EgoNeRF`,
`models.coordinates.YinYangSphericalCoords` (+ `models.envmap.EnvironmentMap`) -- and must leave every other name of the
shadowed modules reachable (`renderer.evaluation`, `models.tensoRF`, `models.tensorBase`, `models.sh`, the other eight
coordinate systems), because `train.py:6,11-15` imports them.  So each shim module loads its reference counterpart from
the reference tree by file location, re-exports everything, and overrides only the names of the hot path.

The reference tree is the first `sys.path` entry after this directory that holds `renderer.py` and `models/EgoNeRF.py`
(the directory `train.py` is run from), or `$EGONERF_REFERENCE`.
"""
import importlib.util
import os
import sys

SHIM_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(SHIM_DIR)

_root = None


def _is_reference(p):
    return (os.path.isfile(os.path.join(p, "renderer.py")) and os.path.isfile(os.path.join(p, "models", "EgoNeRF.py"))
            and os.path.isfile(os.path.join(p, "models", "tensorBase.py")))


def reference_root(required=True):
    """Path of the reference checkout, or None (`required=False`) when there is none: the shim then exposes only the
    names of the B200 path (enough for `from renderer import volume_renderer` / `from models.EgoNeRF import EgoNeRF`),
    and any other name of the shadowed modules raises an ImportError that says why."""
    global _root
    if _root is not None:
        return _root
    cands = []
    if os.environ.get("EGONERF_REFERENCE"):
        cands.append(os.environ["EGONERF_REFERENCE"])
    cands += [p if p else os.getcwd() for p in sys.path]
    for p in cands:
        p = os.path.abspath(p)
        if p != SHIM_DIR and _is_reference(p):
            _root = p
            return p
    if not required:
        return None
    raise ImportError("egonerf_b200 shim: the reference tree (renderer.py, models/EgoNeRF.py) is not on sys.path behind "
                      f"{SHIM_DIR}; run from the reference checkout or set EGONERF_REFERENCE")


def ensure_package_importable():
    """`egonerf_b200` lives next to this directory; make it importable when only shim/ was put on PYTHONPATH."""
    try:
        import egonerf_b200  # noqa: F401
    except ImportError:
        sys.path.append(REPO_ROOT)
        import egonerf_b200  # noqa: F401


def load_reference_module(relpath, alias, package=None):
    """Executes <reference>/<relpath> as module `alias` (registered in sys.modules so that pickles and relative imports
    inside it resolve).  `package` sets `__package__` for files that use `from .x import y`."""
    if alias in sys.modules:
        return sys.modules[alias]
    path = os.path.join(reference_root(), relpath)
    spec = importlib.util.spec_from_file_location(alias, path)
    mod = importlib.util.module_from_spec(spec)
    if package is not None:
        mod.__package__ = package
    sys.modules[alias] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        del sys.modules[alias]
        raise
    return mod


def missing_name_hook(module_name):
    """Module-level __getattr__ for a shim module that could not find its reference counterpart."""
    def __getattr__(name):
        if name.startswith("__") and name.endswith("__"):          # import machinery probes (__path__, __all__, ...)
            raise AttributeError(name)
        raise ImportError(f"egonerf_b200 shim: `{module_name}.{name}` lives in the reference checkout, which is not on "
                          f"sys.path behind {SHIM_DIR} (run from the reference tree or set EGONERF_REFERENCE)")
    return __getattr__


def reexport(mod, namespace):
    """Copies every public name of `mod` (what `from mod import *` and attribute access would see) into `namespace`."""
    for k, v in vars(mod).items():
        if not (k.startswith("__") and k.endswith("__")):
            namespace.setdefault(k, v)
