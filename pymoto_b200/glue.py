"""Device glue for the compliance loop: SIMP interpolation and the compliance inner product.

In the reference these are generic host modules (``MathExpression("xmin + (1-xmin)*inp0^3")``,
pymoto/modules/generic.py:14-140, and ``EinSum("i,i->")``, :143-226).  Here they always run on the GPU (libpmb
kernels / the deterministic device reduction); like every module of this package they accept numpy arrays or CUDA
tensors and return the kind they were given, so a design iteration can stay resident in HBM between the filter, the
assembly and the solve.  (The reference's own MathExpression / EinSum remain usable on numpy Signals next to them.)
"""
from . import _lib
from . import device as dv
from .core import Module


class SIMP(Module):
    """s = xmin + (1 - xmin) * y**p"""

    def __init__(self, xmin=1e-9, p=3):
        if float(p) != int(p) or int(p) < 1:
            raise ValueError(f"pymoto_b200.SIMP evaluates integer penalisation exponents p >= 1 (got {p}); "
                             "use the reference's MathExpression on numpy Signals for a fractional exponent")
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


class Sum(Module):
    """v = sum_i x_i (``EinSum('i->')``, the volume of a density field): deterministic device reduction."""

    _ones = {}

    def __call__(self, x):
        xd = dv.to_device(x).reshape(-1)
        self._n, self._host = xd.numel(), not dv.is_device(x)
        key = (xd.device.index, self._n)
        if key not in Sum._ones:
            import torch

            Sum._ones.clear()  # one cached vector of ones (the current problem size)
            Sum._ones[key] = torch.ones(self._n, dtype=torch.float64, device=xd.device)
        from . import slab

        v = slab.context().comm.allreduce_(dv.dots([(xd, Sum._ones[key])]))[0]  # global sum over the slabs
        return float(v.item()) if self._host else v

    def _sensitivity(self, dvol):
        import torch

        val = float(dvol.item()) if dv.is_device(dvol) else float(dvol)
        d = torch.full((self._n,), val, dtype=torch.float64, device=dv.require_cuda())
        return d.cpu().numpy() if self._host else d


class Scaling(Module):
    """Objective / constraint scaling of a scalar response (pymoto/modules/scaling.py:8-125): objective
    ``y = x * scaling / |x_0|``; constraint ``(x - maxval)/|maxval| * scaling`` or ``(minval - x)/|minval| * scaling`` (or the
    smoothed two-sided form).  Scalar arithmetic only -- works on host floats and on 0-d CUDA tensors alike."""

    def __init__(self, scaling: float = 100.0, minval: float = None, maxval: float = None, minmax_smooth: float = 1e-2):
        self.minval, self.maxval, self.scaling, self.minmax_smooth = minval, maxval, scaling, minmax_smooth
        self.sf = None
        self.reset_scaling()

    def __call__(self, x):
        self._x = x
        if self.sf is None:
            self.sf = self.scaling / abs(float(x))
        if self.minval is not None and self.maxval is not None:
            midp, diff = (self.minval + self.maxval) / 2, (self.maxval - self.minval) / 2
            g = ((x - midp) ** 2 + self.minmax_smooth * diff ** 2) ** 0.5 - ((1 + self.minmax_smooth) * diff ** 2) ** 0.5
        elif self.minval is not None:
            g = (self.minval - x) / (1 if self.minval == 0 else abs(self.minval))
        elif self.maxval is not None:
            g = (x - self.maxval) / (1 if self.maxval == 0 else abs(self.maxval))
        else:
            g = x
        return g * self.sf

    def _sensitivity(self, dy):
        dg = dy * self.sf
        if self.minval is not None and self.maxval is not None:
            midp, diff = (self.minval + self.maxval) / 2, (self.maxval - self.minval) / 2
            return dg * (self._x - midp) / ((self._x - midp) ** 2 + self.minmax_smooth * diff ** 2) ** 0.5
        if self.minval is not None:
            return -dg / (1 if self.minval == 0 else abs(self.minval))
        if self.maxval is not None:
            return dg / (1 if self.maxval == 0 else abs(self.maxval))
        return dg

    def reset_scaling(self):
        self.sf = self.scaling if (self.minval is not None or self.maxval is not None) else None
