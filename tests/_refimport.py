"""Import the UNMODIFIED reference (pyMOTO at /root/reference) for oracle pinning; returns None when absent.

matplotlib is not installed in this image and the reference imports it at module level
(pymoto/common/domain.py:10-11, pymoto/modules/io.py:5-8), so empty stand-in modules are injected first.
Only tests and tests/golden/make_golden.py use this; the GPU box has no /root/reference.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PYMOTO_REFERENCE", "/root/reference")


def import_reference():
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "pymoto")):
        return None
    if "pymoto" in sys.modules:
        return sys.modules["pymoto"]
    try:
        import matplotlib  # noqa: F401
    except ModuleNotFoundError:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        patches = types.ModuleType("matplotlib.patches")
        patches.PathPatch = type("PathPatch", (), {})
        path = types.ModuleType("matplotlib.path")
        path.Path = type("Path", (), {})
        pyplot = types.ModuleType("matplotlib.pyplot")
        mpl.patches, mpl.path, mpl.pyplot = patches, path, pyplot
        sys.modules.update({"matplotlib": mpl, "matplotlib.patches": patches, "matplotlib.path": path,
                            "matplotlib.pyplot": pyplot})
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import pymoto
    except Exception:
        sys.path.remove(REFERENCE_ROOT)
        return None
    return pymoto
