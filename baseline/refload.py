"""Locate and import the UNMODIFIED reference (pyMOTO): /root/reference when it exists (build container), else the
git-ignored install baseline/_ref (travels to the GPU box; made by baseline/install_ref.py).  Returns None when neither is
there.  matplotlib is not in the image and the reference imports it at module level (pymoto/common/domain.py:10-11,
pymoto/modules/io.py:5-8): empty stand-in modules are injected first, nothing in the reference is edited.

Used by tests/ (oracle pinning) and by bench.py's CPU arms only -- never by the product path.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = [os.environ.get("PYMOTO_REFERENCE", "/root/reference"), os.path.join(HERE, "_ref")]


def _stub_matplotlib():
    try:
        import matplotlib  # noqa: F401
        return
    except ModuleNotFoundError:
        pass
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    patches = types.ModuleType("matplotlib.patches")
    patches.PathPatch = type("PathPatch", (), {})
    path = types.ModuleType("matplotlib.path")
    path.Path = type("Path", (), {})
    pyplot = types.ModuleType("matplotlib.pyplot")
    mpl.patches, mpl.path, mpl.pyplot = patches, path, pyplot
    sys.modules.update({"matplotlib": mpl, "matplotlib.patches": patches, "matplotlib.path": path, "matplotlib.pyplot": pyplot})


def reference_root():
    for root in CANDIDATES:
        if root and os.path.isdir(os.path.join(root, "pymoto")):
            return root
    return None


def import_reference():
    if "pymoto" in sys.modules:
        return sys.modules["pymoto"]
    root = reference_root()
    if root is None:
        return None
    _stub_matplotlib()
    sys.path.insert(0, root)
    try:
        import pymoto
    except Exception:
        sys.path.remove(root)
        return None
    return pymoto
