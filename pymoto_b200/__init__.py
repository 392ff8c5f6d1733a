"""pymoto_b200: the per-design-iteration hot path of pyMOTO (SIMP stiffness assembly -> CG + geometric multigrid
solve -> density filter and sensitivities) as hand-written sm_100a CUDA kernels behind pyMOTO's Module /
LinearSolver API.  See DESIGN.md, include/pmb.h and INTEGRATION.md.

    import pymoto_b200 as pmb
    domain = pmb.VoxelDomain(64, 32, 32)
    mgs = pmb.solvers.auto_multigrid(domain)
    K = pmb.AssembleStiffness(domain, bc=bc)(s)
    u = pmb.LinSolve(hermitian=True, solver=pmb.solvers.CG(preconditioner=mgs[0], tol=1e-8))(K, f)

Importing the package does not need a GPU; constructing any module does (there is no CPU fallback).
"""
from .core import Signal, Module, Network, HAVE_PYMOTO
from .domain import VoxelDomain, DomainDefinition
from .matrix import DeviceCSR
from .dyad import DeviceDyad
from .assembly import AssembleGeneral, AssembleStiffness, AssemblePoisson
from .filter import DensityFilter, Filter, FilterConv
from .linalg import LinSolve
from .glue import SIMP, Compliance, Sum, Scaling
from . import solvers
from .optimizers import OC, minimize_oc, MMA, minimize_mma
from .io import WriteToVTI, write_to_vti
from . import slab
from ._lib import PmbError

__all__ = ["Signal", "Module", "Network", "VoxelDomain", "DomainDefinition", "DeviceCSR", "DeviceDyad",
           "AssembleGeneral", "AssembleStiffness", "AssemblePoisson", "DensityFilter", "Filter", "FilterConv", "LinSolve", "SIMP", "Compliance", "Sum", "Scaling", "solvers", "slab", "OC", "minimize_oc", "MMA", "minimize_mma", "WriteToVTI", "write_to_vti",
           "PmbError", "HAVE_PYMOTO"]
