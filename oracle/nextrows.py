"""Oracle restatements for the "next" rows that are built: FilterConv and the OC update
(test infrastructure only).

  FilterConv           pymoto/modules/filter.py:8-220   (padded index array + scipy.signal convolve / correlate)
  OC.step (update)     pymoto/common/optimizers.py:402-437
"""
from numbers import Number

import numpy as np
from scipy.signal import convolve, correlate

from .grid import Grid


class FilterConv:
    def __init__(self, grid: Grid, radius=None, weights=None, xmin_bc="symmetric", xmax_bc="symmetric", ymin_bc="symmetric",
                 ymax_bc="symmetric", zmin_bc="symmetric", zmax_bc="symmetric"):
        self.grid = grid
        if (weights is None) == (radius is None):
            raise ValueError("Only one of arguments 'filter_radius' or 'weights' must be provided.")
        if weights is not None:
            self.weights = np.array(weights, dtype=float)
            while self.weights.ndim < 3:
                self.weights = np.expand_dims(self.weights, axis=-1)
        else:  # filter.py:189-205 (relative units)
            n = [grid.nelx, grid.nely, grid.nelz]
            d = [min(n[a], int(radius - 1e-10)) for a in range(3)]
            cx, cy, cz = np.meshgrid(*[np.arange(-v, v + 1) * 1.0 for v in d], indexing="ij")
            self.weights = np.maximum(0.0, radius - np.sqrt(cx * cx + cy * cy + cz * cz))
            self.weights /= np.sum(self.weights)
        self.pad = [v // 2 for v in self.weights.shape]
        self.n = [grid.nelx, grid.nely, max(grid.nelz, 1)]
        self.overrides = []
        ex, ey, ez = np.meshgrid(*[np.arange(v) for v in self.n], indexing="ij")
        self.el3d_orig = grid.elem_number(ex, ey, ez)
        a = self._pad(self.el3d_orig, xmin_bc, xmax_bc, 0)
        a = self._pad(a, ymin_bc, ymax_bc, 1)
        self.el3d_pad = self._pad(a, zmin_bc, zmax_bc, 2)

    def _pad(self, idx, bc0, bc1, axis):
        p = self.pad[axis]
        pw = lambda lo, hi: [(lo, hi) if a == axis else (0, 0) for a in range(3)]  # noqa: E731
        if bc0 == "wrap" or bc1 == "wrap":
            idx = np.pad(idx, pw(p if bc0 == "wrap" else 0, p if bc1 == "wrap" else 0), mode="wrap")
        padded = [self.n[a] + 2 * self.pad[a] for a in range(3)]
        for side, bc in ((1, bc1), (0, bc0)):  # the reference handles the max edge first
            width = pw(0, p) if side else pw(p, 0)
            if bc in ("edge", "symmetric"):
                idx = np.pad(idx, width, mode=bc)
            elif isinstance(bc, Number):
                idx = np.pad(idx, width, mode="constant", constant_values=0)
                rng = [np.arange(v) for v in padded]
                rng[axis] = (p + self.n[axis] + np.arange(p)) if side else np.arange(p)
                if p > 0:
                    self.overrides.append((tuple(np.meshgrid(*rng, indexing="ij")), bc))
        return idx

    def override_values(self, index, value):
        rng = [self.pad[a] + np.arange(self.n[a]) for a in range(3)]
        ex, ey, ez = np.meshgrid(*rng, indexing="ij")
        self.overrides.append(((ex[index], ey[index], ez[index]), value))

    def __call__(self, x):
        xpad = x[self.el3d_pad]
        for index, value in self.overrides:
            xpad[index] = value
        y = np.zeros_like(x)
        np.add.at(y, self.el3d_orig, convolve(xpad, self.weights, mode="valid"))
        return y

    def sensitivity(self, dy, nel):
        dx3d = correlate(dy[self.el3d_orig], self.weights, mode="full")
        for index, _ in self.overrides:
            dx3d[index] = 0
        dx = np.zeros(nel)
        np.add.at(dx, self.el3d_pad, dx3d)
        return dx


def oc_update(x, dg, move=0.1, xmin=0.0, xmax=1.0, maxvol=None, l1=0.0, l2=100000.0, tol=1e-4):
    """One optimality-criteria design update (optimizers.py:416-435)."""
    if maxvol is None:
        maxvol = np.sum(x) / x.size
    dg = np.minimum(dg, 0)
    lb, ub = np.maximum(xmin, x - move), np.minimum(xmax, x + move)
    xnew = x.copy()
    while l2 - l1 > tol:
        lmid = 0.5 * (l1 + l2)
        xnew[:] = np.clip(x * np.sqrt(-dg / lmid), lb, ub)
        l1, l2 = (lmid, l2) if np.sum(xnew) - maxvol * x.size > 0 else (l1, lmid)
    return xnew
