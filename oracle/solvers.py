"""Oracle: damped Jacobi, geometric multigrid, (block-free) PCG and the LDAS wrapper (test infrastructure only).

Restates, for real symmetric matrices and a single right-hand side (the compliance path):
  DampedJacobi                       pymoto/solvers/iterative.py:21-47
  GeometricMultigrid                 pymoto/solvers/iterative.py:124-256
  CG (+ orth on one column)          pymoto/solvers/iterative.py:259-403
  get_diagonal_indices, LDAWrapper   pymoto/solvers/solvers.py:88-306
  coarsest level = sparse LU         pymoto/solvers/auto_determine.py:109-122, sparse.py:533-550
The sparse products stay on scipy (``csr_matvec`` / ``csc_matvec`` / ``csr_matmat`` / ``splu``) exactly like
the reference.
"""
import numpy as np
import scipy.sparse as sps
from scipy.sparse.linalg import splu

from .grid import Grid


class SparseLU:
    def update(self, A):
        self.lu = splu(sps.csc_matrix(A))

    def solve(self, rhs, x0=None):
        return self.lu.solve(rhs)


class DampedJacobi:
    def __init__(self, w=1.0):
        self.w = w

    def update(self, A):
        self.D = A.diagonal()

    def solve(self, r, x0=None):
        return self.w * (r.T / self.D).T


def prolongation_matrix(fine: Grid, coarse: Grid, ndof):
    """Trilinear prolongation R (nfine x ncoarse), weights 1, 1/2, 1/4, 1/8; no BC awareness
    (iterative.py:178-220).  Built per coarse-node offset like the reference, returned as CSR."""
    rows, cols, vals = [], [], []
    zoffs = (-1, 0, 1) if fine.dim == 3 else (0,)
    for di in (-1, 0, 1):
        ic = np.arange(max(-di, 0), min(coarse.nelx + 1 - di, coarse.nelx + 1))
        for dj in (-1, 0, 1):
            jc = np.arange(max(-dj, 0), min(coarse.nely + 1 - dj, coarse.nely + 1))
            for dk in zoffs:
                kc = np.arange(max(-dk, 0), min(coarse.nelz + 1 - dk, coarse.nelz + 1))
                I, J, K = np.meshgrid(ic, jc, kc, indexing="ij")
                nc = coarse.node_number(I, J, K).ravel()
                nf = fine.node_number(2 * I + di, 2 * J + dj, 2 * K + dk).ravel()
                w = 0.5 ** (abs(di) + abs(dj) + abs(dk))
                for d in range(ndof):
                    rows.append(nf * ndof + d)
                    cols.append(nc * ndof + d)
                    vals.append(np.full(nf.size, w))
    R = sps.coo_matrix(
        (np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
        shape=(ndof * fine.nnodes, ndof * coarse.nnodes),
    )
    return sps.csr_matrix(R)


class GeometricMultigrid:
    def __init__(self, grid: Grid, inner_level=None, smoother=None, smooth_steps=5):
        assert grid.nelx % 2 == 0 and grid.nely % 2 == 0 and grid.nelz % 2 == 0
        self.grid = grid
        self.sub_grid = grid.coarsen()
        self.inner_level = inner_level
        self.smoother = DampedJacobi(w=0.5) if smoother is None else smoother
        self.smooth_steps = smooth_steps
        self.R = None

    def update(self, A):
        if self.R is None:
            ndof = A.shape[0] // self.grid.nnodes
            self.R = prolongation_matrix(self.grid, self.sub_grid, ndof)
        self.A = A
        self.smoother.update(A)
        self.Ac = self.R.T @ A @ self.R  # iterative.py:173
        if self.inner_level is None:
            self.inner_level = SparseLU()
        self.inner_level.update(self.Ac)

    def solve(self, rhs, x0=None):
        # iterative.py:222-256 (trans == "N")
        u = self.smoother.solve(rhs)
        for _ in range(self.smooth_steps - 1):
            r = rhs - self.A @ u
            u += self.smoother.solve(r)
        r = rhs - self.A @ u
        r_c = self.R.T @ r
        u_c = self.inner_level.solve(r_c)
        u += self.R @ u_c
        for _ in range(self.smooth_steps):
            r = rhs - self.A @ u
            u += self.smoother.solve(r)
        return u


def make_gmg_chain(grid: Grid, min_size=8, max_levels=None):
    """Level chain of examples/topology_optimization/ex_compliance_multigrid.py:107-121."""
    mgs = [GeometricMultigrid(grid)]
    while True:
        sub = mgs[-1].sub_grid
        if any(n % 2 != 0 for n in sub.size) or any(sub.size < min_size):
            break
        if max_levels is not None and len(mgs) >= max_levels:
            break
        mgs.append(GeometricMultigrid(sub))
        mgs[-2].inner_level = mgs[-1]
    return mgs


class Identity:
    def update(self, A):
        pass

    def solve(self, r, x0=None):
        return r.copy()


class CG:
    """Single-column restatement of the reference block PCG (iterative.py:295-403)."""

    def __init__(self, preconditioner=None, tol=1e-7, maxit=10000, restart=50):
        self.preconditioner = Identity() if preconditioner is None else preconditioner
        self.tol, self.maxit, self.restart = tol, maxit, restart
        self.iterations = 0
        self.residual = None

    def update(self, A):
        self.A = A
        self.preconditioner.update(A)

    def solve(self, rhs, x0=None):
        A = self.A
        b = rhs
        x = np.zeros_like(rhs) if x0 is None else x0.copy()
        r = b - A @ x
        bnorm = np.linalg.norm(b)
        tval = np.linalg.norm(r) / bnorm
        self.iterations = 0
        self.residual = tval
        if tval <= self.tol:
            return x
        z = self.preconditioner.solve(r)
        p = z / np.sqrt(z @ z)  # orth(z, normalize=True) on one column (iterative.py:259-292)
        for i in range(self.maxit):
            q = A @ p
            pq = p @ q
            alpha = (p @ r) / pq
            x += p * alpha
            if i % self.restart == 0:
                r = b - A @ x
            else:
                r -= q * alpha
            tval = np.linalg.norm(r) / bnorm
            self.iterations = i + 1
            self.residual = tval
            if tval <= self.tol:
                break
            z = self.preconditioner.solve(r)
            beta = -(q @ z) / pq
            p = z + p * beta
        return x


def diagonal_only_rows(A):
    """Rows whose only non-zero is the diagonal, i.e. Dirichlet dofs (solvers.py:88-96)."""
    b = A != 0
    has_diag = b.diagonal()
    nnz_r = np.asarray(b.sum(axis=0)).ravel()
    return np.logical_and(has_diag, nnz_r <= 1)


class LDAWrapper:
    """Linear-dependency-aware wrapper, real symmetric, one rhs (solvers.py:99-306)."""

    def __init__(self, solver, tol=1e-7):
        self.solver, self.tol = solver, tol
        self.x_stored, self.b_stored = [], []
        self.did_solve = False
        self.last_rtol = 0.0

    def update(self, A):
        self.A = A
        diags = diagonal_only_rows(A)
        self.idia = np.flatnonzero(diags)
        self.isel = np.flatnonzero(~diags)
        self.x_stored.clear()
        self.b_stored.clear()
        self.solver.update(A)

    def solve(self, rhs, x0=None):
        A, isel, idia = self.A, self.isel, self.idia
        rhs_loc = np.array(rhs, dtype=float)
        sol = np.zeros_like(rhs_loc)
        sol[idia] = rhs_loc[idia] / A.diagonal()[idia]
        rhs_loc[idia] = 0
        for x, b in zip(self.x_stored, self.b_stored):
            alpha = rhs_loc[isel] @ b / (b @ b)
            rhs_loc[isel] -= alpha * b
            sol[isel] += alpha * x
        bnorm = np.linalg.norm(rhs)
        if bnorm == 0:
            bnorm = 1
        self.last_rtol = np.linalg.norm(A @ sol - rhs) / bnorm
        self.did_solve = self.last_rtol > self.tol
        if self.did_solve:
            x0_loc = None
            if x0 is not None:
                x0_loc = x0.copy()
                x0_loc[idia] = 0
                for x in self.x_stored:
                    beta = x0_loc[isel] @ x / (x @ x)
                    x0_loc[isel] -= beta * x
            xnew = self.solver.solve(rhs_loc, x0_loc)
            sol[isel] += xnew[isel]
            xadd = xnew[isel]
            badd = (A @ xnew)[isel]
            for x, b in zip(self.x_stored, self.b_stored):
                beta = badd @ b / (b @ b)
                badd -= beta * b
                xadd -= beta * x
            bnrm = np.linalg.norm(badd)
            if np.isfinite(bnrm) and bnrm != 0:
                self.x_stored.append(xadd / bnrm)
                self.b_stored.append(badd / bnrm)
        return sol
