"""Device glue for the compliance loop: SIMP interpolation and the compliance inner product.

In the reference these are generic host modules (``MathExpression("xmin + (1-xmin)*inp0^3")``,
pymoto/modules/generic.py:14-140, and ``EinSum("i,i->")``, :143-226).  Here they always run on the GPU (libpmb
kernels / the deterministic device reduction); like every module of this package they accept numpy arrays or CUDA
tensors and return the kind they were given, so a design iteration can stay resident in HBM between the filter, the
assembly and the solve.  (The reference's own MathExpression / EinSum remain usable on numpy Signals next to them.)
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
        yd = dv.to_device(y).reshape(-1)
        self._y = yd
        s = dv.empty(yd.numel())
        _lib.call("pmb_simp", yd.numel(), self.xmin, self.p, dv.ptr(yd), dv.ptr(s), dv.stream())
        return dv.like_input(s, y)

    def _sensitivity(self, ds):
        dsd = dv.to_device(ds).reshape(-1)
        dy = dv.empty(self._y.numel())
        _lib.call("pmb_simp_bwd", self._y.numel(), self.xmin, self.p, dv.ptr(self._y), dv.ptr(dsd), dv.ptr(dy), dv.stream())
        return dv.like_input(dy, ds)


class Compliance(Module):
    """c = u . f (``EinSum('i,i->')``): deterministic device reduction (+ all-reduce over the slabs when distributed)."""

    def __call__(self, u, f):
        from . import slab

        ctx = slab.context()
        if ctx.active and not dv.is_device(u):
            raise TypeError("distributed runs keep nodal vectors on the device")
        self._u, self._f = dv.to_device(u).reshape(-1), dv.to_device(f).reshape(-1)
        self._host = not dv.is_device(u)
        c = ctx.comm.allreduce_(dv.dots([(self._u, self._f)]))[0]
        return float(c.item()) if self._host else c

    def _sensitivity(self, dc):
        dcv = float(dc.item()) if dv.is_device(dc) else float(dc)
        du, df = dv.empty(self._u.numel()), dv.empty(self._u.numel())
        dv.lincomb(du, dcv, self._f)
        dv.lincomb(df, dcv, self._u)
        if self._host:
            return du.cpu().numpy(), df.cpu().numpy()
        return du, df
