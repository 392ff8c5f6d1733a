"""Import the UNMODIFIED reference (pyMOTO) for oracle pinning; returns None when absent.  Thin alias of
baseline/refload.py (which also serves bench.py's CPU arms): /root/reference in the build container, baseline/_ref on the
GPU box."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
from refload import CANDIDATES, import_reference, reference_root  # noqa: E402,F401

REFERENCE_ROOT = CANDIDATES[0]
