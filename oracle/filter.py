"""Oracle: density filter (test infrastructure only).

Restates ``pymoto/modules/filter.py``:
  H_ij = max(0, r - dist(i, j)) on the (2*int(r)+1)^dim window clipped to the domain   filter.py:307-378
  Hs = H.sum(1); y = (H x)/Hs; backward dx = H (dy/Hs)                                 filter.py:241-270
H is kept as a CSC matrix like the reference (``.tocsc()``, filter.py:249) so the product runs through the
same scipy ``csc_matvec`` and accumulates each row in ascending column (= element-number) order.
"""
import numpy as np
import scipy.sparse as sps

from .grid import Grid


def density_filter_matrix(grid: Grid, radius=2.0):
    d = int(radius)
    nx, ny, nz = grid.nelx, grid.nely, max(grid.nelz, 1)
    e = np.arange(grid.nel)
    ix, iy, iz = grid.elem_indices(e)
    rows, cols, vals = [], [], []
    # vectorised over elements per window offset (the reference loops over elements, filter.py:363-368;
    # the resulting COO set is identical, and COO->CSC conversion sorts it)
    for dz in range(-d, d + 1):
        for dy in range(-d, d + 1):
            for dx in range(-d, d + 1):
                jx, jy, jz = ix + dx, iy + dy, iz + dz
                ok = (jx >= 0) & (jx < nx) & (jy >= 0) & (jy < ny) & (jz >= 0) & (jz < nz)
                w = max(0.0, radius - np.sqrt(float(dx * dx + dy * dy + dz * dz)))
                rows.append(e[ok])
                cols.append(grid.elem_number(jx[ok], jy[ok], jz[ok]))
                vals.append(np.full(int(ok.sum()), w))
    H = sps.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(grid.nel, grid.nel))
    return H.tocsc()


class DensityFilter:
    def __init__(self, grid: Grid, radius=2.0, nonpadding=None):
        self.H = density_filter_matrix(grid, radius)
        self.Hs = np.asarray(self.H.sum(1))  # (nel, 1), filter.py:251
        if nonpadding is not None:  # filter.py:253-255
            inds = ~np.isin(np.arange(len(self.Hs)), nonpadding)
            self.Hs[inds] = np.max(self.Hs)

    def __call__(self, x):
        return np.asarray(self.H @ x[np.newaxis].T / self.Hs)[:, 0]  # filter.py:266-267

    def sensitivity(self, dy):
        return np.asarray(self.H @ (dy[np.newaxis].T / self.Hs))[:, 0]  # filter.py:269-270
