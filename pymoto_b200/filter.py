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
        # the kernel's window is trimmed to the offsets that carry weight (an integer radius r leaves the outermost shell of the
        # (2r+1)^dim window at exactly 0: 27 of 125 slots at r = 2); w * x = +-0 terms never change a sum, the bits stay
        nzw = w > 0
        self._dk = dk = int(max(np.abs(dx[nzw]).max(), np.abs(dy[nzw]).max(), np.abs(dz[nzw]).max())) if nzw.any() else 0
        keep = (np.abs(dx) <= dk) & (np.abs(dy) <= dk) & (np.abs(dz) <= dk)
        self._wtab = dv.to_device(np.ascontiguousarray(w[keep]))
        # row sums Hs = H @ 1 (filter.py:251) by the same stencil with unit input
        self.Hs = dv.empty(self.nel)
        self._apply(None, None, self.Hs)
        if nonpadding is not None:  # filter.py:253-255
            if ctx.active:
                raise NotImplementedError("pymoto_b200.DensityFilter(nonpadding=...) is not distributed over z-slabs "
                                          "(global element numbers and a global max of Hs are involved)")
            keep = torch.zeros(self.nel, dtype=torch.bool, device=self.Hs.device)
            keep[dv.to_device(np.asarray(nonpadding).ravel(), torch.int64)] = True
            self.Hs = torch.where(keep, self.Hs, self.Hs.max())

    def _apply(self, inp, hs, out):
        if inp is not None and self._ctx.active:
            self._xbuf[self._pad: self._pad + self.nel] = inp
            self._ctx.comm.exchange(self._xbuf, self._pad, self.nel, self._lay, width=self.d)
            inp = self._xbuf[self._pad: self._pad + self.nel]
        _lib.call("pmb_filter_apply", self.grid, self._e0, self.nlayers, self._dk, dv.ptr(self._wtab), dv.ptr(inp), dv.ptr(hs),
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


class FilterConv(Module):
    r"""Density filter as a padded convolution, :math:`y = W \ast x` (pymoto/modules/filter.py:8-220).

    Same constructor as the reference: either ``radius`` (cone kernel, normalised to unit sum) or explicit ``weights``;
    per-face boundary treatment ``"symmetric"`` (default), ``"edge"``, ``"wrap"`` or a constant value.  The padded
    index array of the reference is replaced by per-axis index maps (built with the same ``np.pad`` sequence) and
    the convolution by a direct stencil on the GPU; ``override_values`` / ``override_padded_values`` are kept.
    """

    def __init__(self, domain, radius: float = None, relative_units: bool = True, weights=None, xmin_bc="symmetric",
                 xmax_bc="symmetric", ymin_bc="symmetric", ymax_bc="symmetric", zmin_bc="symmetric", zmax_bc="symmetric"):
        dv.require_cuda()
        self.domain = domain
        self.weights = None
        if (weights is None and radius is None) or (weights is not None and radius is not None):
            raise ValueError("Only one of arguments 'filter_radius' or 'weights' must be provided.")
        elif weights is not None:
            self.weights = np.array(weights, dtype=float)
            while self.weights.ndim < 3:
                self.weights = np.expand_dims(self.weights, axis=-1)
            for i in range(self.weights.ndim):
                assert self.weights.shape[i] % 2 == 1, "Size of weights must be uneven"
        else:
            self.set_filter_radius(radius, relative_units)
        nx, ny, nz = grid_dims(domain)
        self.pad_sizes = [v // 2 for v in self.weights.shape]
        maps = [self._axis_map(n, self.pad_sizes[a], bc0, bc1)
                for a, (n, (bc0, bc1)) in enumerate(zip((nx, ny, max(nz, 1)), [(xmin_bc, xmax_bc), (ymin_bc, ymax_bc), (zmin_bc, zmax_bc)]))]
        # slab decomposition: this rank filters its element layers [e0, e1).  The "source" field of the kernels is then the
        # rank's layers plus pz halo layers on each side (filled from the z-neighbours), and the z map is the window
        # [e0, e1 + 2 pz) of the global one re-based to that extended field; global faces keep their boundary treatment.
        self._ctx = ctx = slab.context(nz)
        self._halo = 0
        if ctx.active:
            pz = self.pad_sizes[2]
            e0, e1 = ctx.part.elem_layers(0)
            if "wrap" in (zmin_bc, zmax_bc):
                raise NotImplementedError("FilterConv: z wrap-around boundaries are not available under a slab decomposition")
            if e1 - e0 < pz:
                raise ValueError(f"FilterConv kernel needs at least {pz} element layers per rank (have {e1 - e0})")
            mz, cz = maps[2]
            win = mz[e0: e1 + 2 * pz].copy()
            loc = np.where(win >= 0, win - (e0 - pz), -1)
            assert np.all((loc < e1 - e0 + 2 * pz)), "boundary treatment reaches beyond the neighbouring slab"
            maps[2] = (loc, cz[e0: e1 + 2 * pz].copy())
            self._halo, self._lay, self._nown = pz, nx * ny, (e1 - e0)
            self._n = (nx, ny, e1 - e0 + 2 * pz)          # source extent seen by the gather / scatter kernels
            self._nout = (nx, ny, e1 - e0)                # layers this rank produces
            self._p = (nx + 2 * self.pad_sizes[0], ny + 2 * self.pad_sizes[1], e1 - e0 + 2 * pz)
            self.nel = nx * ny * (e1 - e0)
        else:
            self._n = self._nout = (nx, ny, max(nz, 1))
            self.nel = nx * ny * max(nz, 1)
            self._p = tuple(n + 2 * p for n, p in zip(self._n, self.pad_sizes))
        self.overrides = []  # user overrides: (flat padded indices (device int64), value)
        self._map = [dv.to_device(m[0].astype(np.int32), torch.int32) for m in maps]
        self._cval = [dv.to_device(m[1]) for m in maps]
        # inverse maps (which padded positions read a given source index), CSR form, for the backward scatter
        self._inv = []
        for a, (m, _) in enumerate(maps):
            order = np.argsort(m, kind="stable")
            order = order[m[order] >= 0]
            counts = np.bincount(m[m >= 0], minlength=self._n[a])
            ptr = np.zeros(len(counts) + 1, dtype=np.int32)
            np.cumsum(counts, out=ptr[1:])
            self._inv.append((dv.to_device(ptr, torch.int32), dv.to_device(order.astype(np.int32), torch.int32)))
        self._upload_weights()

    def _upload_weights(self):
        w = np.ascontiguousarray(self.weights, dtype=np.float64)
        # device layout is (z, y, x) with x fastest; scipy's convolve flips the kernel, correlate does not
        self._w_bwd = dv.to_device(np.ascontiguousarray(w.transpose(2, 1, 0)).ravel())
        self._w_fwd = dv.to_device(np.ascontiguousarray(w[::-1, ::-1, ::-1].transpose(2, 1, 0)).ravel())

    @staticmethod
    def _axis_map(n, pad, bc0, bc1):
        """Source index of every padded position along one axis (-1 = constant padding) and the constant values;
        the reference's sequence: wrap sides first, then the max edge, then the min edge (filter.py:99-160)."""
        from numbers import Number

        idx = np.arange(n)
        cval = np.zeros(n + 2 * pad)
        if pad == 0:
            return idx, cval
        wrap = (pad if bc0 == "wrap" else 0, pad if bc1 == "wrap" else 0)
        a = np.pad(idx, wrap, mode="wrap") if (wrap[0] or wrap[1]) else idx
        if bc1 == "edge":
            a = np.pad(a, (0, pad), mode="edge")
        elif bc1 == "symmetric":
            a = np.pad(a, (0, pad), mode="symmetric")
        elif isinstance(bc1, Number):
            a = np.pad(a, (0, pad), mode="constant", constant_values=-1)
            cval[pad + n:] = bc1
        elif bc1 != "wrap":
            raise ValueError(f"Unknown boundary condition {bc1!r}")
        if bc0 == "edge":
            a = np.pad(a, (pad, 0), mode="edge")
        elif bc0 == "symmetric":
            a = np.pad(a, (pad, 0), mode="symmetric")
        elif isinstance(bc0, Number):
            a = np.pad(a, (pad, 0), mode="constant", constant_values=-1)
            cval[:pad] = bc0
        elif bc0 != "wrap":
            raise ValueError(f"Unknown boundary condition {bc0!r}")
        assert a.size == n + 2 * pad
        return a, cval

    def set_filter_radius(self, radius: float, relative_units: bool = True):
        """Cone kernel max(0, r - dist), normalised (filter.py:189-205)."""
        if relative_units:
            dx, dy, dz = 1.0, 1.0, 1.0
        else:
            dx, dy, dz = self.domain.element_size
        nx, ny, nz = grid_dims(self.domain)
        delemx = min(nx, int((radius - 1e-10 * dx) / dx))
        delemy = min(ny, int((radius - 1e-10 * dy) / dy))
        delemz = min(nz, int((radius - 1e-10 * dz) / dz))
        cx, cy, cz = np.meshgrid(np.arange(-delemx, delemx + 1) * dx, np.arange(-delemy, delemy + 1) * dy,
                                 np.arange(-delemz, delemz + 1) * dz, indexing="ij")
        self.weights = np.maximum(0.0, radius - np.sqrt(cx * cx + cy * cy + cz * cz))
        self.weights /= np.sum(self.weights)  # volume preserving
        if hasattr(self, "_w_fwd"):
            self._upload_weights()

    # ---- user overrides of padded / domain values (filter.py:167-180)
    def override_padded_values(self, index, value):
        if self._ctx.active:
            raise NotImplementedError("FilterConv overrides address the global padded grid: not available under a slab decomposition")
        ix, iy, iz = (np.asarray(i) for i in index)
        if ix.size == 0:
            return
        flat = (np.broadcast_arrays(ix, iy, iz)[2] * self._p[1] + np.broadcast_arrays(ix, iy, iz)[1]) * self._p[0] \
            + np.broadcast_arrays(ix, iy, iz)[0]
        self.overrides.append((dv.to_device(flat.ravel().astype(np.int64), torch.int64), float(value)))

    def override_values(self, index, value):
        xr = self.pad_sizes[0] + np.arange(self._n[0])
        yr = self.pad_sizes[1] + np.arange(self._n[1])
        zr = self.pad_sizes[2] + np.arange(self._n[2])
        ex, ey, ez = np.meshgrid(xr, yr, zr, indexing="ij")
        self.override_padded_values((ex[index], ey[index], ez[index]), value)

    def _extended(self, xd):
        """This rank's layers with pz halo layers from the z-neighbours on each side (halos at the global faces are never read)."""
        h = self._halo * self._lay
        ext = dv.zeros(self.nel + 2 * h)
        ext[h: h + self.nel] = xd
        if h:
            self._ctx.comm.exchange(ext, h, self.nel, self._lay, width=self._halo)
        return ext

    def get_padded_vector(self, x):
        """Padded field on the device, flat in (z, y, x) order with x fastest."""
        xd = dv.to_device(x).reshape(-1)
        if self._ctx.active:
            xd = self._extended(xd)
        xpad = dv.empty(self._p[0] * self._p[1] * self._p[2])
        _lib.call("pmb_pad_gather", *self._n, *self._p, dv.ptr(self._map[0]), dv.ptr(self._map[1]), dv.ptr(self._map[2]),
                  dv.ptr(self._cval[0]), dv.ptr(self._cval[1]), dv.ptr(self._cval[2]), dv.ptr(xd), dv.ptr(xpad), dv.stream())
        for idx, value in self.overrides:
            xpad[idx] = value
        return xpad

    def __call__(self, x):
        n = x.numel() if isinstance(x, torch.Tensor) else np.size(x)
        if n != self.nel:
            raise ValueError(f"Input vector wrong size ({n}), must be equal to #nel ({self.nel})")
        xpad = self.get_padded_vector(x)
        y = dv.empty(self.nel)
        k = self.weights.shape
        _lib.call("pmb_stencil_corr", *self._p, dv.ptr(xpad), *self._nout, dv.ptr(y), k[0], k[1], k[2], dv.ptr(self._w_fwd),
                  0, 0, 0, dv.stream())
        return dv.like_input(y, x)

    def _sensitivity(self, dfdv):
        dy = dv.to_device(dfdv).reshape(-1)
        k = self.weights.shape
        dxpad = dv.empty(self._p[0] * self._p[1] * self._p[2])
        _lib.call("pmb_stencil_corr", *self._nout, dv.ptr(dy), *self._p, dv.ptr(dxpad), k[0], k[1], k[2], dv.ptr(self._w_bwd),
                  k[0] - 1, k[1] - 1, k[2] - 1, dv.stream())
        for idx, _ in self.overrides:
            dxpad[idx] = 0.0
        dx = dv.empty(self._n[0] * self._n[1] * self._n[2])
        _lib.call("pmb_pad_scatter", *self._n, *self._p, dv.ptr(self._inv[0][0]), dv.ptr(self._inv[0][1]),
                  dv.ptr(self._inv[1][0]), dv.ptr(self._inv[1][1]), dv.ptr(self._inv[2][0]), dv.ptr(self._inv[2][1]),
                  dv.ptr(dxpad), dv.ptr(dx), dv.stream())
        if self._ctx.active and self._halo:
            # contributions that landed in my halo layers belong to the z-neighbours: send them back and add what the
            # neighbours collected for my outermost layers (reverse of the forward halo exchange)
            h = self._halo * self._lay
            t = dv.zeros(4 * h)  # [lower halo | for the rank below | for the rank above | upper halo]
            t[h: 2 * h] = dx[:h]
            t[2 * h: 3 * h] = dx[h + self.nel:]
            self._ctx.comm.exchange(t, h, 2 * h, self._lay, width=self._halo)
            own = dx[h: h + self.nel].clone()
            part = self._ctx.part
            if part.lower is not None:
                own[:h] += t[:h]
            if part.upper is not None:
                own[self.nel - h:] += t[3 * h:]
            dx = own
        return dv.like_input(dx, dfdv)
