"""
TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Imports the UNMODIFIED reference (`/root/reference/bundle_adjust/*.py`) inside
this build container so that
  * the numpy restatement in `oracle/ba_oracle.py` can be validated against it,
  * golden fixtures under `tests/golden/` can be generated from it
    (`tests/golden/make_golden.py`).

`/root/reference` does not exist on the GPU box, so nothing that runs there may
call `load_reference()`; `reference_available()` is the guard.

The reference package cannot be imported as-is: `bundle_adjust/__init__.py`
pulls in rpcm / rasterio / srtm4 / pyproj / utm / matplotlib, none of which is
installed.  We register empty stand-in modules for those
names and a bare namespace package whose `__path__` points at the reference
tree, which is enough for ba_core, ba_params, ba_rotate, cam_utils, geo_utils.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SBA_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "rasterio", "rasterio.errors",
    "rpcm", "pyproj", "utm", "srtm4", "shapely", "shapely.geometry", "ad",
]


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "bundle_adjust", "ba_core.py"))


def load_reference():
    """Returns a namespace with the reference's hot-path modules (ba_core, ba_params, ...)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "bundle_adjust" not in sys.modules or not getattr(sys.modules["bundle_adjust"], "_sba_stub", False):
        for name in _STUBS:
            if name not in sys.modules:
                sys.modules[name] = types.ModuleType(name)
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        sys.modules["rasterio"].errors = sys.modules["rasterio.errors"]
        if not hasattr(sys.modules["rasterio.errors"], "NotGeoreferencedWarning"):
            sys.modules["rasterio.errors"].NotGeoreferencedWarning = type(
                "NotGeoreferencedWarning", (UserWarning,), {})
        pkg = types.ModuleType("bundle_adjust")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "bundle_adjust")]
        pkg._sba_stub = True
        sys.modules["bundle_adjust"] = pkg
    # the reference's ba_rpcfit builds its output model with `rpcm.RPCModel(dict)`; rpcm is not installed, so the
    # stub module hands out the oracle's look-alike (validated against the reference's compiled C, see rpc_oracle.py)
    from . import rpc_oracle
    sys.modules["rpcm"].RPCModel = rpc_oracle.RPCModel
    old = sys.dont_write_bytecode
    sys.dont_write_bytecode = True  # the reference tree is read-only
    try:
        from bundle_adjust import ba_core, ba_params, ba_rotate, ba_rpcfit, cam_utils, geo_utils
    finally:
        sys.dont_write_bytecode = old
    ns = types.SimpleNamespace(ba_core=ba_core, ba_params=ba_params, ba_rotate=ba_rotate, ba_rpcfit=ba_rpcfit,
                               cam_utils=cam_utils, geo_utils=geo_utils)
    return ns
