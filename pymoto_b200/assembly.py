"""Assembly modules with the reference's constructor signatures, producing a DeviceCSR on the GPU.

Mirrors pymoto/modules/assembly.py: ``AssembleGeneral`` (:18-315), ``AssembleStiffness`` (:407-463),
``AssemblePoisson`` (:523-560).  The reference precomputes a CSR pattern and a scatter map (:100-230) and runs
``np.add.at`` per call (:255-275); here the pattern is closed-form and the values are written by
``pmb_assemble`` (bit-identical values, see pmb_assembly.cu).  Backward: ``pmb_assemble_sens`` (:298-315).
"""
import numpy as np
import torch

from . import _lib
from . import device as dv
from .core import Module
from .domain import grid_dims
from .dyad import DeviceDyad
from .matrix import DeviceCSR, ElemGenerator, make_grid
from . import slab


def _strain_displacement(dN_dx):
    """B in Voigt order [xx, yy, zz, yz, zx, xy] (3-D) / [xx, yy, xy] (2-D); cf. get_B, assembly.py:318-370."""
    dim, nn = dN_dx.shape
    B = np.zeros((dim * (dim + 1) // 2, nn * dim), dtype=dN_dx.dtype)
    for a in range(nn):
        c = a * dim
        if dim == 2:
            B[0, c], B[1, c + 1] = dN_dx[0, a], dN_dx[1, a]
            B[2, c], B[2, c + 1] = dN_dx[1, a], dN_dx[0, a]
        elif dim == 3:
            B[0, c], B[1, c + 1], B[2, c + 2] = dN_dx[0, a], dN_dx[1, a], dN_dx[2, a]
            B[3, c + 1], B[3, c + 2] = dN_dx[2, a], dN_dx[1, a]
            B[4, c], B[4, c + 2] = dN_dx[2, a], dN_dx[0, a]
            B[5, c], B[5, c + 1] = dN_dx[1, a], dN_dx[0, a]
        else:
            raise ValueError(f"Number of dimensions ({dim}) must be 2 or 3")
    return B


def _elasticity_matrix(E, nu, mode):
    """cf. get_D, assembly.py:373-404."""
    mu = E / (2 * (1 + nu))
    lam = (E * nu) / ((1 + nu) * (1 - 2 * nu))
    c1 = 2 * mu + lam
    if "strain" in mode:
        return np.array([[c1, lam, 0], [lam, c1, 0], [0, 0, mu]])
    if "stress" in mode:
        return E / (1 - nu * nu) * np.array([[1, nu, 0], [nu, 1, 0], [0, 0, (1 - nu) / 2]])
    if "3d" in mode:
        D = np.zeros((6, 6))
        D[:3, :3] = lam
        D[np.arange(3), np.arange(3)] = c1
        D[np.arange(3, 6), np.arange(3, 6)] = mu
        return D
    raise ValueError("Only for plane-stress, plane-strain, or 3d")


def _gauss_points(domain):
    siz = domain.element_size
    for n in domain.node_numbering:
        yield np.asarray(n) * (siz / 2) / np.sqrt(3)


class AssembleGeneral(Module):
    r"""``A = sum_e sum_i x_{i,e} A_{i,e}`` on a structured grid, Dirichlet rows/columns zeroed with ``bcdiagval`` on the
    diagonal, ``add_constant`` added afterwards (assembly.py:18-296).

    Inputs: one scaling vector of size ``(nel)`` per element matrix (numpy arrays or CUDA tensors).  Output: :class:`DeviceCSR`.

    One element matrix and no constant is the hot path: values bit-identical to the reference's, matrix-free finest-level operator
    and direct level-1 Galerkin build attached.  Several element matrices are assembled one by one and summed (the reference adds
    the scaled element matrices before scattering, so the values agree to rounding, not bit for bit) and, like ``add_constant``
    (a scipy sparse matrix whose entries lie inside the grid's 27-point node stencil), are used as assembled values on every
    multigrid level.
    """

    def __init__(self, domain, element_matrix, bc=None, bcdiagval=None, matrix_type=None, add_constant=None,
                 reuse_sparsity: bool = True):
        elmats = list(element_matrix) if isinstance(element_matrix, (list, tuple)) else [element_matrix]
        self.elmat = [np.asarray(m) for m in elmats]
        self.nmat = len(self.elmat)
        if self.nmat < 1:
            raise ValueError("No or invalid element-matrix is given")
        Ke = self.elmat[0]
        if any(np.iscomplexobj(m) for m in self.elmat):
            raise TypeError("complex element matrices are not supported by the B200 hot path (real FP64 only)")
        elemnodes = domain.elemnodes
        if Ke.shape[0] % elemnodes != 0:
            raise ValueError("Number of rows in element matrix should be a multiple of the number of nodes per element")
        if Ke.shape[1] % elemnodes != 0:
            raise ValueError("Number of cols in element matrix should be a multiple of the number of nodes per element")
        for m in self.elmat[1:]:
            if m.shape != Ke.shape:
                raise ValueError(f"Element matrices must be the same shape {m.shape} != {Ke.shape}")
        self.mdof = Ke.shape[0] // elemnodes
        self.ndof = Ke.shape[1] // elemnodes
        if self.mdof != self.ndof:
            raise NotImplementedError("pymoto_b200.AssembleGeneral supports square matrices only")
        if not 1 <= self.ndof <= 3:
            raise NotImplementedError("pymoto_b200 supports 1..3 dofs per node")
        if matrix_type is not None and getattr(matrix_type, "__name__", "") not in ("csr_matrix", "csr_array", "DeviceCSR"):
            raise NotImplementedError("pymoto_b200 assembles CSR (DeviceCSR) only")
        self.domain = domain
        self.reuse_sparsity = reuse_sparsity  # the pattern is closed-form: nothing to precompute

        dv.require_cuda()
        nx, ny, nz = grid_dims(domain)
        # slab decomposition (one GPU: the whole grid).  Inputs / outputs of the module are this rank's element layers.
        self._ctx = ctx = slab.context(nz)
        k0, k1 = ctx.part.planes(0) if ctx.active else (0, nz + 1)
        e0, e1 = ctx.part.elem_layers(0) if ctx.active else (0, max(nz, 1))
        self.grid = make_grid(nx, ny, nz, self.ndof, k0, k1 - k0)
        self._lay = nx * ny  # elements per layer
        self.nel = self._lay * (e1 - e0)
        plane = (nx + 1) * (ny + 1) * self.ndof
        self.m = self.n = plane * (k1 - k0)
        self._Ke_hosts = [np.ascontiguousarray(m, dtype=np.float64).ravel().copy() for m in self.elmat]
        self._Ke_devs = [dv.to_device(k) for k in self._Ke_hosts]
        self._Ke_host, self._Ke_dev = self._Ke_hosts[0], self._Ke_devs[0]
        self.add_constant = add_constant
        if add_constant is not None:
            import scipy.sparse as sps

            if not sps.issparse(add_constant):
                raise TypeError("pymoto_b200.AssembleGeneral: add_constant must be a scipy sparse matrix")
            if ctx.active:
                raise NotImplementedError("pymoto_b200.AssembleGeneral: add_constant is not distributed over z-slabs")
            if add_constant.shape != (self.m, self.n):
                raise ValueError(f"add_constant has shape {add_constant.shape}, the assembled matrix {(self.m, self.n)}")
        self._const_vals = None  # add_constant on the positions of the stencil-CSR pattern (device), built on the first call

        self.bc = None
        self.bcdiagval = bcdiagval
        self._bcmask = self._bcmask_buf = None
        if bc is not None:
            self.bc = np.asarray(bc).ravel()  # GLOBAL dof numbers, the same on every rank
            if bcdiagval is None:
                self.bcdiagval = np.max(sum(self.elmat[1:], self.elmat[0]))  # assembly.py:94-98
            # local mask over planes [k0-1, k1+1): the assembly kernel also looks at the column dofs in the halo planes
            lo = (k0 - 1) * plane
            mask = np.zeros((k1 - k0 + 2) * plane, dtype=np.uint8)
            sel = self.bc[(self.bc >= max(lo, 0)) & (self.bc < (k1 + 1) * plane)]
            mask[sel - lo] = 1
            self._bcmask_buf = dv.to_device(mask, torch.uint8)
            self._bcmask = self._bcmask_buf[plane: plane + self.n]
        self._xbufs = [None] * self.nmat
        self._mat = None
        self._scratch = None

    def _stage_scaling(self, i, xscale):
        """Private copy of scaling vector ``i`` (it also generates the matrix-free finest-level operator until the next call)
        with room for the element layers below my first node plane, fetched from the rank below (two layers: the direct
        level-1 Galerkin build reads the children of the coarse element layer below the slab)."""
        n = xscale.numel() if isinstance(xscale, torch.Tensor) else np.size(xscale)
        if n != self.nel:
            raise ValueError(f"Input vector wrong size ({n}), must be equal to #nel ({self.nel})")
        x = dv.to_device(xscale).reshape(-1)
        if self._xbufs[i] is None:
            self._xbufs[i] = dv.zeros(self.nel + 2 * self._lay)
        buf = self._xbufs[i]
        buf[2 * self._lay:] = x
        if self._ctx.active:
            self._ctx.comm.exchange(buf, 2 * self._lay, self.nel, self._lay, lower=True, upper=False, width=2)
        return buf[2 * self._lay:]

    def _constant_values(self, mat):
        """``add_constant`` as values on the stencil-CSR pattern of ``mat`` (built once; entries outside the pattern refuse)."""
        if self._const_vals is None:
            import scipy.sparse as sps

            c = sps.coo_matrix(self.add_constant)
            if np.iscomplexobj(c.data):
                raise TypeError("complex add_constant is not supported by the B200 hot path (real FP64 only)")
            indptr, indices = (t.cpu().numpy() for t in (mat.indptr, mat.indices))
            slot = sps.csr_matrix((np.arange(1, mat.nnz + 1, dtype=np.int64), indices, indptr), shape=mat.shape)
            pos = np.asarray(slot[c.row, c.col]).ravel() - 1
            outside = (pos < 0) & (c.data != 0)
            if np.any(outside):
                raise NotImplementedError(f"pymoto_b200.AssembleGeneral: add_constant has {int(outside.sum())} entries outside the "
                                          "27-point node stencil of the grid")
            vals = np.zeros(mat.nnz)
            np.add.at(vals, pos[pos >= 0], np.asarray(c.data, dtype=np.float64)[pos >= 0])
            self._const_vals = dv.to_device(vals)
        return self._const_vals

    def __call__(self, *xscale):
        if len(xscale) != self.nmat:
            raise ValueError(f"One scaling vector must be given for each element matrix ({self.nmat})")
        self._x_on_device = [dv.is_device(x) for x in xscale]
        xs = [self._stage_scaling(i, x) for i, x in enumerate(xscale)]
        x = xs[0]
        # a fresh value buffer each call would cost 8*nnz bytes of allocation per design iteration; the matrix
        # object is reused and its cached row statistics dropped (consumers re-read it on every update()).
        if self._mat is None:
            self._mat = DeviceCSR(self.grid, bc_mask=self._bcmask, comm=self._ctx.comm, level=0)
        mat = self._mat
        bcdiag = float(self.bcdiagval if self.bcdiagval is not None else 0.0)
        diag, nnz_off = mat.rowstats_buffers()  # filled by the assembly kernel while the rows are on chip
        _lib.call("pmb_assemble", self.grid, self._Ke_host.ctypes.data, dv.ptr(x), dv.ptr(self._bcmask), bcdiag, dv.ptr(mat._buf),
                  dv.ptr(diag), dv.ptr(nnz_off), dv.stream())
        mat.invalidate()
        if self.nmat == 1 and self.add_constant is None:  # the hot path: row statistics and the matrix-free description come along
            mat._diag, mat._nnz_off = diag, nnz_off
            if mat.generator is None:
                mat.generator = ElemGenerator(self.grid, self._Ke_host, x, self._bcmask, bcdiag, bc=self.bc)
            else:
                mat.generator.retarget(x)
            mat.autotune_matrix_free()
            return mat
        # several element matrices (assembly.py:245-253) and / or a constant (:294-295): the other terms are assembled with a
        # zero Dirichlet diagonal into a scratch buffer and added; the operator is applied from its assembled values
        for i in range(1, self.nmat):
            if self._scratch is None:
                self._scratch = dv.empty(mat.nnz + 2)
            _lib.call("pmb_assemble", self.grid, self._Ke_hosts[i].ctypes.data, dv.ptr(xs[i]), dv.ptr(self._bcmask), 0.0,
                      dv.ptr(self._scratch), None, None, dv.stream())
            dv.lincomb(mat.data, 1.0, mat.data, 1.0, self._scratch[: mat.nnz])
        if self.add_constant is not None:
            dv.lincomb(mat.data, 1.0, mat.data, 1.0, self._constant_values(mat))
        return mat

    def _sensitivity(self, dgdmat):
        if dgdmat is None or getattr(dgdmat, "size", 1) <= 0:
            return [None] * self.nmat
        if not isinstance(dgdmat, DeviceDyad):
            raise TypeError("pymoto_b200.AssembleGeneral back-propagates a DeviceDyad (from pymoto_b200.LinSolve)")
        dxs = [dv.zeros(self.nel) for _ in range(self.nmat)]
        first = True
        mat = self._mat
        for u, v in zip(dgdmat.u, dgdmat.v):
            if mat is not None and mat.comm is not None:  # element layer e needs node planes e and e+1
                u, v = mat.operand(u), mat.operand(v)
                mat.exchange(u, lower=False, upper=True)
                mat.exchange(v, lower=False, upper=True)
            for ke, dx in zip(self._Ke_devs, dxs):
                _lib.call("pmb_assemble_sens", self.grid, dv.ptr(ke), dv.ptr(u), dv.ptr(v), dv.ptr(self._bcmask), dv.ptr(dx),
                          0 if first else 1, dv.stream())
            first = False
        on_dev = getattr(self, "_x_on_device", [False] * self.nmat)
        return [dx if dev else dx.cpu().numpy() for dx, dev in zip(dxs, on_dev)]


class AssembleStiffness(AssembleGeneral):
    r"""Stiffness matrix ``K = sum_e x_e K_e`` for quad4 / hex8 linear elasticity (assembly.py:407-463)."""

    def __init__(self, domain, *args, e_modulus: float = 1.0, poisson_ratio: float = 0.3, plane="strain", **kwargs):
        self.E, self.nu = e_modulus, poisson_ratio
        D = _elasticity_matrix(self.E, self.nu, "3d" if domain.dim == 3 else plane.lower())
        nd = (2 ** domain.dim) * domain.dim
        self.stiffness_element = np.zeros((nd, nd))
        siz = domain.element_size
        w = np.prod(siz[: domain.dim] / 2)
        if domain.dim == 2:
            w *= siz[2]  # thickness
        for pos in _gauss_points(domain):
            B = _strain_displacement(domain.eval_shape_fun_der(pos))
            self.stiffness_element += w * B.T @ D @ B
        super().__init__(domain, self.stiffness_element, *args, **kwargs)


class AssemblePoisson(AssembleGeneral):
    r"""Scalar diffusion matrix ``P = sum_e x_e P_e`` (thermal conduction etc., assembly.py:523-560)."""

    def __init__(self, domain, *args, material_property: float = 1.0, **kwargs):
        self.material_property = material_property
        self.poisson_element = np.zeros((domain.elemnodes, domain.elemnodes))
        siz = domain.element_size
        w = np.prod(siz[: domain.dim] / 2)
        if domain.dim != 3:
            self.material_property *= siz[domain.dim:]
        for pos in _gauss_points(domain):
            Bn = domain.eval_shape_fun_der(pos)
            self.poisson_element += w * self.material_property * Bn.T @ Bn
        super().__init__(domain, self.poisson_element, *args, **kwargs)
