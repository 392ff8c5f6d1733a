"""Device-resident glue for the compliance loop: SIMP interpolation and the compliance inner product.

In the reference these are generic host modules (``MathExpression("xmin + (1-xmin)*inp0^3")``,
pymoto/modules/generic.py:14-140, and ``EinSum("i,i->")``, :143-226).  They accept numpy arrays (computed with
numpy, exactly the reference's expressions) or CUDA tensors (computed by libpmb kernels / a device reduction), so
a design iteration can stay resident in HBM between the filter, the assembly and the solve.
"""
import numpy as np

from . import _lib
from . import device as dv
from .core import Module


class SIMP(Module):
    """s = xmin + (1 - xmin) * y**p"""

    def __init__(self, xmin=1e-9, p=3):
        self.xmin, self.p = float(xmin), int(p)

    def __call__(self, y):
        self._y = y
        if not dv.is_device(y):
            return self.xmin + (1.0 - self.xmin) * y ** self.p
        s = dv.empty(y.numel())
        _lib.call("pmb_simp", y.numel(), self.xmin, self.p, dv.ptr(y), dv.ptr(s), dv.stream())
        return s

    def _sensitivity(self, ds):
        y = self._y
        if not dv.is_device(y):
            return ds * (self.p * (1.0 - self.xmin) * y ** (self.p - 1))
        dsd = dv.to_device(ds)
        dy = dv.empty(y.numel())
        _lib.call("pmb_simp_bwd", y.numel(), self.xmin, self.p, dv.ptr(y), dv.ptr(dsd), dv.ptr(dy), dv.stream())
        return dy


class Compliance(Module):
    """c = u . f (``EinSum('i,i->')``); on device the reduction is the deterministic pmb_dots kernel."""

    def __call__(self, u, f):
        self._u, self._f = u, f
        from . import slab

        ctx = slab.context()
        if not dv.is_device(u):
            if ctx.active:
                raise TypeError("distributed runs keep nodal vectors on the device")
            return np.asarray(u) @ (f.cpu().numpy() if dv.is_device(f) else np.asarray(f))
        return ctx.comm.allreduce_(dv.dots([(u, dv.to_device(f))]))[0]

    def _sensitivity(self, dc):
        u, f = self._u, self._f
        if not dv.is_device(u):
            dc = float(dc)
            return dc * np.asarray(f), dc * np.asarray(u)
        fd = dv.to_device(f)
        if dv.is_device(dc):
            return dc * fd, dc * u
        return float(dc) * fd, float(dc) * u
