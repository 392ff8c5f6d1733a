"""Density filter with the reference's signature, evaluated as a stencil on the GPU.

Mirrors ``Filter`` / ``DensityFilter`` of pymoto/modules/filter.py:223-378: ``y = (H x)/Hs`` with
``H_ij = max(0, r - dist(i, j))`` on the ``(2*int(r)+1)^dim`` window clipped to the domain, ``Hs = H.sum(1)``
(optionally overridden through ``nonpadding``), backward ``dx = H (dy / Hs)``.  The matrix H (125 stored
entries per row at r = 2, built by a Python loop over all elements in the reference) is never formed.
"""
import numpy as np
import torch

from . import _lib
from . import device as dv
from .core import Module
from .domain import grid_dims
from .matrix import make_grid
from . import slab


class DensityFilter(Module):
    def __init__(self, domain, radius=2.0, nonpadding=None):
        dv.require_cuda()
        self.domain = domain
        self.radius = radius
        nx, ny, nz = grid_dims(domain)
        self.grid = make_grid(nx, ny, nz, 1)
        d = int(radius)  # window half-width, filter.py:314
        self.d = d
        # slab decomposition: this rank filters its own element layers [e0, e1) and reads d halo layers on each side
        self._ctx = ctx = slab.context(nz)
        self._e0, e1 = ctx.part.elem_layers(0) if ctx.active else (0, max(nz, 1))
        self.nlayers = e1 - self._e0
        self._lay = nx * ny
        self.nel = self._lay * self.nlayers
        if ctx.active and self.nlayers < d:
            raise ValueError(f"filter radius {radius} needs at least {d} element layers per rank (have {self.nlayers})")
        self._pad = self._lay * d if ctx.active else 0
        self._xbuf = dv.zeros(self.nel + 2 * self._pad) if ctx.active else None
        # cone weights exactly as the reference computes them (integer offsets -> sqrt -> max), filter.py:371-375
        rng = np.arange(-d, d + 1)
        if nz > 0:
            dz, dy, dx = np.meshgrid(rng, rng, rng, indexing="ij")
        else:
            dy, dx = np.meshgrid(rng, rng, indexing="ij")
            dz = np.zeros_like(dx)
        w = np.maximum(0.0, radius - np.sqrt(dx * dx + dy * dy + dz * dz))
        self._wtab = dv.to_device(np.ascontiguousarray(w.ravel()))
        # row sums Hs = H @ 1 (filter.py:251) by the same stencil with unit input
        self.Hs = dv.empty(self.nel)
        self._apply(None, None, self.Hs)
        if nonpadding is not None:  # filter.py:253-255
            keep = torch.zeros(self.nel, dtype=torch.bool, device=self.Hs.device)
            keep[dv.to_device(np.asarray(nonpadding).ravel(), torch.int64)] = True
            self.Hs = torch.where(keep, self.Hs, self.Hs.max())

    def _apply(self, inp, hs, out):
        if inp is not None and self._ctx.active:
            self._xbuf[self._pad: self._pad + self.nel] = inp
            self._ctx.comm.exchange(self._xbuf, self._pad, self.nel, self._lay, width=self.d)
            inp = self._xbuf[self._pad: self._pad + self.nel]
        _lib.call("pmb_filter_apply", self.grid, self._e0, self.nlayers, self.d, dv.ptr(self._wtab), dv.ptr(inp), dv.ptr(hs),
                  dv.ptr(out), dv.stream())
        return out

    def _check(self, x):
        n = x.numel() if isinstance(x, torch.Tensor) else np.size(x)
        if n != self.nel:
            raise ValueError(f"Input vector wrong size ({n}), must be equal to #nel ({self.nel})")

    def __call__(self, x):
        self._check(x)
        xd = dv.to_device(x).reshape(-1)
        y = self._apply(xd, self.Hs, dv.empty(self.nel))
        return dv.like_input(y, x)

    def _sensitivity(self, dfdy):
        self._check(dfdy)
        dy = dv.to_device(dfdy).reshape(-1)
        t = dv.empty(self.nel)
        _lib.call("pmb_vec_div", self.nel, dv.ptr(dy), dv.ptr(self.Hs), dv.ptr(t), dv.stream())
        dx = self._apply(t, None, dv.empty(self.nel))
        return dv.like_input(dx, dfdy)


Filter = DensityFilter
