"""LinSolve module with the reference's signature, solving on the GPU.

Mirrors pymoto/modules/linalg.py:121-219: the solver is always wrapped in the linear-dependency-aware
``LDAWrapper`` with ``tol = 5 * solver.tol`` (:185-189), updated with the new matrix every call (:192) and
warm-started from the previous solution (:195); the backward pass solves the adjoint system through the same
wrapper (:199-201) and returns the matrix sensitivity as a rank-1 dyad ``(-lam, u)`` and ``db = lam`` (:204-219).
"""
import numpy as np
import torch

from . import device as dv
from .core import Module
from .dyad import DeviceDyad
from .matrix import DeviceCSR
from .solvers import CG, LDAWrapper, LinearSolver, auto_multigrid
from .domain import VoxelDomain


class LinSolve(Module):
    use_lda_solver = True

    def __init__(self, dep_tol: float = 1e-5, hermitian: bool = None, symmetric: bool = None,
                 positive_definite: bool = None, solver: LinearSolver = None):
        self.dep_tol = dep_tol
        self.ishermitian = hermitian
        self.issymmetric = symmetric
        self.ispositivedefinite = positive_definite
        self.solver = solver
        self.u = None  # solution storage (host-visible, same kind as rhs)
        self._u_dev = None

    def __call__(self, mat, rhs):
        if not isinstance(mat, DeviceCSR):
            raise TypeError("pymoto_b200.LinSolve solves a DeviceCSR from pymoto_b200.AssembleGeneral/Stiffness/Poisson; "
                            f"got {type(mat).__name__} (no CPU fallback)")
        if np.iscomplexobj(rhs) if not isinstance(rhs, torch.Tensor) else rhs.is_complex():
            raise TypeError("Complex right-hand-side for a real-valued sparse matrix is not supported.")
        self.issparse, self.iscomplex = True, False
        if self.ishermitian is None:
            self.ishermitian = True  # assembled from one symmetric-by-construction element matrix
        if self.solver is None:
            # the stand-in for auto_determine_solver on this path: CG preconditioned by the geometric multigrid
            # chain of ex_compliance_multigrid.py:107-121
            g = mat.grid
            mgs = auto_multigrid(VoxelDomain(g.nx, g.ny, g.nz))
            self.solver = CG(preconditioner=mgs[0], tol=1e-8)
        if not isinstance(self.solver, LDAWrapper) and self.use_lda_solver:
            kw = dict(hermitian=self.ishermitian, symmetric=self.issymmetric)
            if hasattr(self.solver, "tol"):
                kw["tol"] = self.solver.tol * 5
            self.solver = LDAWrapper(self.solver, **kw)
        self.solver.update(mat)
        self._rhs_on_device = dv.is_device(rhs)
        self._u_dev = self.solver.solve(dv.to_device(rhs), x0=self._u_dev)
        self.u = self._u_dev if self._rhs_on_device else self._u_dev.cpu().numpy()
        return self.u

    def _sensitivity(self, dfdv):
        lam = self.solver.solve(dv.to_device(dfdv), trans="T")
        mat = self.solver.A if hasattr(self.solver, "A") else None
        if self._u_dev.ndim == 1 and mat is not None and mat.comm is not None:
            neg = mat.new_vec()  # halo-padded: the element sensitivity kernel reads one plane above the slab
            dv.lincomb(neg, -1.0, lam)
            return DeviceDyad(neg, self._u_dev), (lam if self._rhs_on_device else lam.cpu().numpy())
        if self._u_dev.ndim > 1:
            dmat = DeviceDyad([-lam[:, i].contiguous() for i in range(lam.shape[1])],
                              [self._u_dev[:, i].contiguous() for i in range(lam.shape[1])])
        else:
            dmat = DeviceDyad(-lam, self._u_dev)
        db = lam if self._rhs_on_device else lam.cpu().numpy()
        return dmat, db
