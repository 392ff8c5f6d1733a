"""Oracle: one compliance design iteration end to end (test infrastructure only).

Chain of SURVEY.md 3.1 / examples/topology_optimization/ex_compliance_multigrid.py:76-150 with DensityFilter:
  y = filter(x); s = xmin + (1-xmin) y^3; K = assemble(s); u = LDAS(CG(GMG)).solve(K, f) warm-started;
  c = f.u; backward: du = f; lam = LDAS.solve(du) (no CG: projection on the stored pair);
  ds_e = (-lam)[e]^T Ke u[e]; dy = ds * 3(1-xmin) y^2; dx = filter^T(dy).
"""
import numpy as np

from .grid import Grid
from . import assembly as oasm
from . import filter as oflt
from . import solvers as osol


def cantilever(grid: Grid):
    """Cantilever: all dofs clamped on face i=0, unit +z (3-D) / +y (2-D) load on the line i=nx, mid height
    (examples/topology_optimization/ex_compliance.py:60-69 style, as SURVEY.md 8d states)."""
    ndof = grid.dim
    nodes = grid.nodes3d()
    bn = nodes[0, :, :].ravel()
    bc = (bn[:, None] * ndof + np.arange(ndof)[None, :]).ravel()
    f = np.zeros(grid.nnodes * ndof)
    if grid.dim == 3:
        ln = nodes[grid.nelx, :, grid.nelz // 2].ravel()
        f[ln * ndof + 2] = 1.0
    else:
        ln = nodes[grid.nelx, grid.nely // 2].ravel()
        f[ln * ndof + 1] = 1.0
    return ndof, np.sort(bc), f


def mbb3d(grid: Grid):
    """Synthetic 3-D half-MBB: u_x=0 on face i=0, u_y=0 on face j=0, u_z=0 on edge (i=nx,k=0),
    unit -z line load on edge (i=0, k=nz)."""
    ndof = 3
    nodes = grid.nodes3d()
    bc = np.concatenate([
        nodes[0, :, :].ravel() * 3 + 0,
        nodes[:, 0, :].ravel() * 3 + 1,
        nodes[grid.nelx, :, 0].ravel() * 3 + 2,
    ])
    f = np.zeros(grid.nnodes * 3)
    f[nodes[0, :, grid.nelz].ravel() * 3 + 2] = -1.0
    return ndof, np.unique(bc), f


def heatsink(grid: Grid):
    """Thermal: T=0 on a centred patch of face i=0, unit heat load on all nodes with i>=1
    (ex_compliance_multigrid.py:60-63 3-D thermal branch with the centred patch of ex_compliance.py:64)."""
    ndof = 1
    nodes = grid.nodes3d()
    ny, nz = grid.nely, grid.nelz
    bc = nodes[0, ny // 4:(ny + 1) - ny // 4, nz // 4:(nz + 1) - nz // 4].ravel()
    f = np.zeros(grid.nnodes)
    f[nodes[1:, :, :].ravel()] = 1.0
    return ndof, np.sort(bc), f


class ComplianceProblem:
    def __init__(self, grid: Grid, kind="cantilever", radius=2.0, xmin=1e-9, tol=1e-8, solver="gmg", min_size=8,
                 max_levels=None):
        self.grid, self.xmin = grid, xmin
        if kind == "cantilever":
            self.ndof, self.bc, self.f = cantilever(grid)
        elif kind == "mbb3d":
            self.ndof, self.bc, self.f = mbb3d(grid)
        elif kind == "heatsink":
            self.ndof, self.bc, self.f = heatsink(grid)
        else:
            raise ValueError(kind)
        Ke = oasm.poisson_element(grid) if kind == "heatsink" else oasm.stiffness_element(grid)
        self.filt = oflt.DensityFilter(grid, radius)
        self.asm = oasm.Assembler(grid, Ke, bc=self.bc)
        if solver == "gmg":
            self.mgs = osol.make_gmg_chain(grid, min_size=min_size, max_levels=max_levels)
            self.cg = osol.CG(self.mgs[0], tol=tol)
            self.solver = osol.LDAWrapper(self.cg, tol=5 * tol)  # linalg.py:185-189
        else:
            self.cg = None
            self.solver = osol.LDAWrapper(osol.SparseLU())
        self.u = None

    def response(self, x):
        self.y = self.filt(x)
        self.s = self.xmin + (1.0 - self.xmin) * self.y ** 3
        self.K = self.asm(self.s)
        self.solver.update(self.K)
        self.u = self.solver.solve(self.f, x0=self.u)  # warm start, linalg.py:195
        self.c = float(self.u @ self.f)
        return self.c

    def sensitivity(self):
        lam = self.solver.solve(self.f)  # dc/du = f; adjoint via LDAS (linalg.py:199-201)
        ds = self.asm.sensitivity(-lam, self.u)
        dy = ds * (3.0 * (1.0 - self.xmin) * self.y ** 2)
        return self.filt.sensitivity(dy)
