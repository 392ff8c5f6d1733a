#!/usr/bin/env python
"""Install the UNMODIFIED reference (pyMOTO, pure Python) into baseline/_ref (git-ignored, travels to the GPU box).

1. `pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference`
   (fails in this image: the build backend `hatchling` is not installed and there is no network);
2. fallback: what that wheel install would have produced for a pure-Python package -- the package directory copied
   verbatim (no file is edited; `diff -r /root/reference/pymoto baseline/_ref/pymoto` is empty).
The reference's hard dependencies numpy / scipy / sympy are in the image; matplotlib is not, bench.py stubs it the same
way tests/_refimport.py does.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("PYMOTO_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def main():
    if not os.path.isdir(os.path.join(SRC, "pymoto")):
        print(f"reference not found at {SRC}; keeping {DST} as is")
        return 0
    shutil.rmtree(DST, ignore_errors=True)
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links", "/opt/wheelhouse",
           "--no-deps", "--target", DST, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    how = "pip"
    if r.returncode != 0 or not os.path.isdir(os.path.join(DST, "pymoto")):
        how = "copy (pip failed: " + (r.stderr.strip().splitlines() or ["?"])[-1] + ")"
        os.makedirs(DST, exist_ok=True)
        shutil.copytree(os.path.join(SRC, "pymoto"), os.path.join(DST, "pymoto"),
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    open(os.path.join(DST, "INSTALL_NOTE.txt"), "w").write(f"installed from {SRC} by baseline/install_ref.py: {how}\n")
    print(f"baseline/_ref ready ({how})")
    return 0


if __name__ == "__main__":
    sys.exit(main())
