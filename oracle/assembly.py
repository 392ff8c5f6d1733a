"""Oracle: element matrices, CSR pattern, SIMP-scaled assembly and its sensitivity (test infrastructure only).

Restates ``pymoto/modules/assembly.py``:
  element stiffness / Poisson matrix by 2^dim Gauss points      assembly.py:318-404, 438-463, 546-558
  sorted-unique CSR pattern + scatter map ("datamap")            assembly.py:100-206
  Dirichlet handling (zero rows/cols, diag = bcdiagval)          assembly.py:208-230, 270-272
  values: np.add.at(data, datamap, Ke.ravel()*x[:,None])         assembly.py:255-275
  sensitivity dx_e = sum_k u_k[dofs_e]^T Ke v_k[dofs_e]          assembly.py:298-315, dyadcarrier.py:286-414
"""
import numpy as np
import scipy.sparse as sps

from .grid import Grid


# ---------------------------------------------------------------- element matrices
def strain_displacement(dN):
    """B matrix in Voigt order [xx, yy, zz, yz, zx, xy] (assembly.py:318-370)."""
    dim, nn = dN.shape
    nstrain = dim * (dim + 1) // 2
    B = np.zeros((nstrain, nn * dim))
    for a in range(nn):
        c = a * dim
        if dim == 2:
            B[0, c] = dN[0, a]
            B[1, c + 1] = dN[1, a]
            B[2, c], B[2, c + 1] = dN[1, a], dN[0, a]
        else:
            B[0, c] = dN[0, a]
            B[1, c + 1] = dN[1, a]
            B[2, c + 2] = dN[2, a]
            B[3, c + 1], B[3, c + 2] = dN[2, a], dN[1, a]
            B[4, c], B[4, c + 2] = dN[2, a], dN[0, a]
            B[5, c], B[5, c + 1] = dN[1, a], dN[0, a]
    return B


def constitutive(E, nu, mode):
    """Isotropic linear-elastic D (assembly.py:373-404)."""
    mu = E / (2 * (1 + nu))
    lam = (E * nu) / ((1 + nu) * (1 - 2 * nu))
    c1 = 2 * mu + lam
    if "strain" in mode:
        return np.array([[c1, lam, 0], [lam, c1, 0], [0, 0, mu]])
    if "stress" in mode:
        a = E / (1 - nu * nu)
        return a * np.array([[1, nu, 0], [nu, 1, 0], [0, 0, (1 - nu) / 2]])
    D = np.zeros((6, 6))
    D[:3, :3] = lam
    D[np.arange(3), np.arange(3)] = c1
    D[np.arange(3, 6), np.arange(3, 6)] = mu
    return D


def stiffness_element(grid: Grid, e_modulus=1.0, poisson_ratio=0.3, plane="strain"):
    """hex8 / quad4 element stiffness (assembly.py:438-463): sum over 2^dim Gauss points of w*B^T D B,
    accumulated in the same order (node_numbering order) and with the same expression grouping."""
    D = constitutive(e_modulus, poisson_ratio, "3d" if grid.dim == 3 else plane.lower())
    siz = grid.element_size
    w = np.prod(siz[: grid.dim] / 2)
    if grid.dim == 2:
        w *= siz[2]
    nd = grid.elemnodes * grid.dim
    Ke = np.zeros((nd, nd))
    sg3 = np.where(grid.node_signs == 0, 0, grid.node_signs).astype(float)
    for a in range(grid.elemnodes):
        pos = sg3[a] * (siz / 2) / np.sqrt(3)
        B = strain_displacement(grid.shape_fun_der(pos))
        Ke += w * B.T @ D @ B
    return Ke


def poisson_element(grid: Grid, material_property=1.0):
    """Scalar conduction element matrix (assembly.py:546-558)."""
    siz = grid.element_size
    w = np.prod(siz[: grid.dim] / 2)
    mp = material_property
    if grid.dim != 3:
        mp = mp * siz[grid.dim:]
    Pe = np.zeros((grid.elemnodes, grid.elemnodes))
    sg3 = grid.node_signs.astype(float)
    for a in range(grid.elemnodes):
        pos = sg3[a] * (siz / 2) / np.sqrt(3)
        Bn = grid.shape_fun_der(pos)
        Pe += w * mp * Bn.T @ Bn
    return Pe


# ---------------------------------------------------------------- pattern
def pattern_unique(grid: Grid, ndof):
    """CSR pattern as the sorted set of unique (row, col) pairs over all element dof pairs.

    This is the definition the reference's 'manual' construction implements (assembly.py:130-206; its
    'unique' variant :116-128 is literally this).  O(nel*(nn*ndof)^2) memory: small grids only.
    Returns indptr, indices (int64) and datamap (nel*(nn*ndof)^2,) in (e, a, b) order.
    """
    dc = grid.dofconn(ndof)
    m = dc.shape[1]
    rows = np.repeat(dc, m, axis=1).ravel()
    cols = np.tile(dc, (1, m)).ravel()
    n = grid.nnodes * ndof
    key = rows.astype(np.int64) * n + cols
    ukey, datamap = np.unique(key, return_inverse=True)
    urows = ukey // n
    indices = ukey % n
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(urows, minlength=n), out=indptr[1:])
    return indptr, indices, datamap.ravel()


def _cnt_prefix(M):
    """Neighbour count in [i-1, i+1] clipped to [0, M-1], and its exclusive prefix sum (size M+1)."""
    i = np.arange(M)
    cnt = 3 - (i == 0) - (i == M - 1)
    pre = np.zeros(M + 1, dtype=np.int64)
    np.cumsum(cnt, out=pre[1:])
    return cnt.astype(np.int64), pre


def pattern_closed_form(grid: Grid, ndof):
    """Same pattern from the 27-/9-point block-stencil closed form.  Scales to big grids.
    Validated against :func:`pattern_unique` and against the reference in tests/test_oracle.py."""
    NX, NY, NZ = grid.nelx + 1, grid.nely + 1, grid.nelz + 1
    cx, px = _cnt_prefix(NX)
    cy, py = _cnt_prefix(NY)
    cz, pz = _cnt_prefix(NZ)
    n = np.arange(grid.nnodes)
    i, j, k = grid.node_indices(n)
    cnt = cx[i] * cy[j] * cz[k]
    rowlen = np.repeat(cnt * ndof, ndof)
    indptr = np.zeros(grid.nnodes * ndof + 1, dtype=np.int64)
    np.cumsum(rowlen, out=indptr[1:])
    # neighbours of every node in ascending node number (z-major, x fastest)
    idx_chunks = []
    # build per node via broadcasting over the 27 offsets, masking invalid ones
    offs = [(dk, dj, di) for dk in (-1, 0, 1) for dj in (-1, 0, 1) for di in (-1, 0, 1)]
    ii = i[:, None] + np.array([o[2] for o in offs])[None, :]
    jj = j[:, None] + np.array([o[1] for o in offs])[None, :]
    kk = k[:, None] + np.array([o[0] for o in offs])[None, :]
    valid = (ii >= 0) & (ii < NX) & (jj >= 0) & (jj < NY) & (kk >= 0) & (kk < NZ)
    nb = grid.node_number(ii, jj, kk)
    # (node, d, nbr, cd) order
    cols = nb[:, None, :, None] * ndof + np.arange(ndof)[None, None, None, :]
    cols = np.broadcast_to(cols, (grid.nnodes, ndof, 27, ndof))
    vmask = np.broadcast_to(valid[:, None, :, None], cols.shape)
    idx_chunks = cols[vmask]
    return indptr, idx_chunks.astype(np.int64)


def datamap_closed_form(grid: Grid, ndof, indptr):
    """Scatter targets for (e, a, b) in the order of Ke.ravel() per element (assembly.py:186-206)."""
    NX, NY, NZ = grid.nelx + 1, grid.nely + 1, grid.nelz + 1
    cx, _ = _cnt_prefix(NX)
    cy, _ = _cnt_prefix(NY)
    e = np.arange(grid.nel)
    ei, ej, ek = grid.elem_indices(e)
    nn = grid.elemnodes
    out = np.empty((grid.nel, nn * ndof, nn * ndof), dtype=np.int64)
    for a in range(nn):
        ai, aj, ak = ei + (a & 1), ej + ((a >> 1) & 1), ek + ((a >> 2) & 1)
        rown = grid.node_number(ai, aj, ak)
        ilo, jlo, klo = np.maximum(ai - 1, 0), np.maximum(aj - 1, 0), np.maximum(ak - 1, 0)
        cxa, cya = cx[ai], cy[aj]
        for b in range(nn):
            bi, bj, bk = ei + (b & 1), ej + ((b >> 1) & 1), ek + ((b >> 2) & 1)
            nbr = ((bk - klo) * cya + (bj - jlo)) * cxa + (bi - ilo)
            for d in range(ndof):
                base = indptr[rown * ndof + d] + nbr * ndof
                for c in range(ndof):
                    out[:, a * ndof + d, b * ndof + c] = base + c
    return out.reshape(-1)


class Assembler:
    """Oracle restatement of ``AssembleGeneral`` for square element matrices (one, or a list with one scaling vector each,
    assembly.py:245-253) and an optional constant matrix added after the boundary conditions (:294-295), CSR output."""

    def __init__(self, grid: Grid, element_matrix, bc=None, bcdiagval=None, closed_form=True, add_constant=None):
        self.grid = grid
        self.Kes = [np.asarray(m, dtype=float) for m in (element_matrix if isinstance(element_matrix, (list, tuple)) else [element_matrix])]
        self.Ke = self.Kes[0]
        self.add_constant = add_constant
        self.ndof = self.Ke.shape[0] // grid.elemnodes
        self.n = grid.nnodes * self.ndof
        self.bc = None if bc is None else np.asarray(bc).ravel()
        self.bcdiagval = bcdiagval
        if self.bc is not None and bcdiagval is None:
            self.bcdiagval = np.max(sum(self.Kes[1:], self.Kes[0]))  # assembly.py:94-98 (maximum of the SUM of the matrices)
        if closed_form:
            self.indptr, self.indices = pattern_closed_form(grid, self.ndof)
            self.datamap = datamap_closed_form(grid, self.ndof, self.indptr)
        else:
            self.indptr, self.indices, self.datamap = pattern_unique(grid, self.ndof)
        if self.bc is not None:
            # assembly.py:208-230: entries whose row or column is constrained are redirected to one dump slot
            # (the first bc diagonal), which is overwritten with bcdiagval afterwards.
            rows = np.repeat(np.arange(self.n), np.diff(self.indptr))
            isbc = np.zeros(self.n, dtype=bool)
            isbc[self.bc] = True
            bad = (isbc[rows] | isbc[self.indices])[self.datamap]
            diagpos = np.flatnonzero(rows == self.indices)  # one per row, ascending row
            self.bcadd = diagpos[self.bc]
            self.datamap = self.datamap.copy()
            self.datamap[bad] = self.bcadd[0]

    def __call__(self, *xs):
        assert len(xs) == len(self.Kes)  # assembly.py:241-242
        scaled = None
        for x, Ke in zip(xs, self.Kes):  # assembly.py:250-257: the scaled element matrices are added BEFORE the scatter
            term = (Ke.ravel()[None, :] * np.asarray(x).ravel()[:, None]).ravel()
            scaled = term if scaled is None else scaled + term
        data = np.zeros(self.indices.size)
        np.add.at(data, self.datamap, scaled)  # assembly.py:267-268
        if self.bc is not None:
            data[self.bcadd] = self.bcdiagval  # assembly.py:270-272
        mat = sps.csr_matrix((data, self.indices, self.indptr), shape=(self.n, self.n))  # assembly.py:275
        if self.add_constant is not None:
            mat = mat + self.add_constant  # assembly.py:294-295
        return mat

    def sensitivity(self, u, v):
        """dx_e = u[dofs_e]^T Ke v[dofs_e] with u, v zeroed at bc (assembly.py:301-303, 311-314)."""
        u = np.array(u, dtype=float)
        v = np.array(v, dtype=float)
        if self.bc is not None:
            u[self.bc] = 0.0
            v[self.bc] = 0.0
        dc = self.grid.dofconn(self.ndof)
        dx = [np.einsum("Ai,ij,Aj->A", u[dc], Ke, v[dc]) for Ke in self.Kes]
        return dx[0] if len(dx) == 1 else dx
