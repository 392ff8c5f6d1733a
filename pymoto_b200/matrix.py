"""DeviceCSR: the assembled system matrix as it lives in HBM.

It is the reference's CSR matrix (``csr_matrix((data, indices, indptr))``, pymoto/modules/assembly.py:275) with
the ``data`` array resident on the GPU.  On a structured voxel grid ``indptr``/``indices`` are a closed form of
the grid, so they are only materialised (bit-exactly, int32 when nnz fits like scipy's downcast) when a user asks
for them: ``.indptr``, ``.indices`` or ``.tocsr()``.  The solver kernels stream ``data`` alone.

With a slab decomposition (pymoto_b200/slab.py) the object holds the rows of this rank's node planes; vectors it
multiplies are views into buffers padded by one halo plane on each side (:meth:`new_vec`), refreshed from the
neighbours by :meth:`exchange` before every operator application.
"""
import numpy as np
import torch

from . import _lib
from . import device as dv


def make_grid(nx, ny, nz, ndof, kz0=0, nzl=None):
    return _lib.Grid(int(nx), int(ny), int(nz), int(ndof), int(kz0), int(nz + 1 if nzl is None else nzl))


class DeviceCSR:
    # Apply finest-level operators matrix-free (from the element scaling vector, pmb_elem.cu) when the assembly module
    # attached its generator; the assembled values stay the source of truth for everything else.  Set to False to
    # stream the CSR values on every level.
    matrix_free = True

    def __init__(self, grid: _lib.Grid, data: torch.Tensor = None, bc_mask: torch.Tensor = None, comm=None, level=0):
        dv.require_cuda()
        self.grid = grid
        self.comm = comm if (comm is not None and comm.active and (grid.kz0 != 0 or grid.nzl != grid.nz + 1)) else None
        self.level = level
        self.nnz = _lib.query("pmb_nnz", grid)
        self.n = _lib.query("pmb_nrows", grid)
        self.shape = (self.n, self.n)  # rows owned by this rank (= the whole matrix on one GPU)
        self.ndim = 2
        self.dtype = np.dtype(np.float64)
        self.plane = (grid.nx + 1) * (grid.ny + 1) * grid.ndof  # dofs per node plane
        if data is None:
            data = dv.empty(self.nnz + 2)  # +2: 16-byte slack read by the 128-bit bulk copies
            data[self.nnz:] = 0.0
        assert data.is_cuda and data.dtype == torch.float64 and data.numel() >= self.nnz + 2
        assert data.data_ptr() % 16 == 0
        self._buf = data
        self.bc_mask = bc_mask  # uint8 per dof or None (set by the assembly module; informational)
        self._indptr = self._indices = None
        self._diag = self._nnz_off = None
        self.generator = None  # dict(ke=host ndarray, s=device tensor, mask=device uint8 or None, bcdiag=float)

    # ---- values
    @property
    def data(self):
        return self._buf[: self.nnz]

    def invalidate(self):
        """Call after the values changed in place."""
        self._diag = self._nnz_off = None

    # ---- vectors this operator can be applied to (halo-padded), halo refresh, global dot products
    def new_vec(self, zero=False):
        """Owned entries of a fresh vector whose storage carries one halo node plane on each side."""
        base = (dv.zeros if zero else dv.empty)(self.n + 2 * self.plane)
        if not zero and self.comm is None:
            pass  # halos are never read on one GPU (the stencil is clipped at the grid faces)
        elif not zero:
            base[: self.plane] = 0.0
            base[self.plane + self.n:] = 0.0
        return base[self.plane: self.plane + self.n]

    def exchange(self, vec, lower=True, upper=True):
        """Refresh the halo planes of a vector made by :meth:`new_vec` from the neighbouring ranks."""
        if self.comm is None:
            return
        base = vec._base if vec._base is not None else vec
        off = vec.storage_offset()
        if off < self.plane or base.numel() < off + self.n + self.plane or vec.numel() != self.n:
            raise _lib.PmbError("distributed operator input must come from DeviceCSR.new_vec() (halo-padded storage)")
        self.comm.exchange(base, off, self.n, self.plane, lower=lower, upper=upper)

    def operand(self, x):
        """``x`` itself if it can be fed to :meth:`apply` (always on one GPU; halo-padded storage when distributed),
        else a padded copy."""
        if self.comm is None:
            return x
        base = x._base
        if base is not None and x.storage_offset() >= self.plane and base.numel() >= x.storage_offset() + self.n + self.plane:
            return x
        xp = self.new_vec()
        xp.copy_(x)
        return xp

    def dots(self, pairs):
        """Global dot products (local deterministic reduction + sum all-reduce over the slabs)."""
        d = dv.dots(pairs)
        if self.comm is not None:
            self.comm.allreduce_(d)
        return d

    # ---- row statistics: diagonal + number of non-zero off-diagonals, one pass over the values
    def rowstats(self):
        if self._diag is None:
            self._diag = dv.empty(self.n)
            self._nnz_off = dv.empty(self.n, torch.int32)
            _lib.call("pmb_rowstats", self.grid, dv.ptr(self._buf), dv.ptr(self._diag), dv.ptr(self._nnz_off), dv.stream())
        return self._diag, self._nnz_off

    def diagonal_device(self):
        return self.rowstats()[0]

    def diagonal(self):
        return self.diagonal_device().cpu().numpy()

    # ---- products
    def apply(self, mode, x, y, b=None, diag=None, w=0.0, dotv=None, dot_out=None):
        """Raw kernel call on device tensors: y = A x | b - A x | x + w (b - A x)/diag, optional fused dots."""
        gen = self.generator if DeviceCSR.matrix_free else None
        ws = None
        if dot_out is not None:
            ws = dv.workspace().spmv_ws(_lib.query("pmb_spmv_ws_doubles" if gen is None else "pmb_elem_ws_doubles", self.grid))
        if self.comm is not None:
            self.exchange(x)
        if gen is None:
            _lib.call("pmb_spmv", self.grid, mode, dv.ptr(self._buf), dv.ptr(x), dv.ptr(b), dv.ptr(diag), float(w), dv.ptr(y),
                      dv.ptr(dotv), dv.ptr(dot_out), dv.ptr(ws), dv.stream())
        else:
            _lib.call("pmb_elem_spmv", self.grid, mode, gen["ke"].ctypes.data, dv.ptr(gen["s"]), dv.ptr(gen["mask"]),
                      float(gen["bcdiag"]), dv.ptr(x), dv.ptr(b), dv.ptr(diag), float(w), dv.ptr(y), dv.ptr(dotv),
                      dv.ptr(dot_out), dv.ptr(ws), dv.stream())
        if dot_out is not None and self.comm is not None:
            self.comm.allreduce_(dot_out)
        return y

    def matvec_device(self, x, out=None):
        y = dv.empty(self.n) if out is None else out
        return self.apply(_lib.SPMV, self.operand(x), y)

    def __matmul__(self, x):
        xd = dv.to_device(x)
        if xd.ndim == 1:
            return dv.like_input(self.matvec_device(xd), x)
        cols = [self.matvec_device(xd[:, i].contiguous()) for i in range(xd.shape[1])]
        return dv.like_input(torch.stack(cols, dim=1), x)

    dot = __matmul__

    # ---- export (parity / interop only; never used by the solver path)
    def _pattern(self):
        if self._indptr is None:
            g = self.grid
            if g.kz0 != 0 or g.nzl != g.nz + 1:
                raise _lib.PmbError("CSR export needs the whole grid on one rank")
            bits = 32 if self.nnz < 2 ** 31 - 1 else 64
            tdt = torch.int32 if bits == 32 else torch.int64
            self._indptr = dv.empty(self.n + 1, tdt)
            self._indices = dv.empty(self.nnz, tdt)
            _lib.call("pmb_csr_pattern", g, dv.ptr(self._indptr), dv.ptr(self._indices), bits, dv.stream())
        return self._indptr, self._indices

    @property
    def indptr(self):
        return self._pattern()[0]

    @property
    def indices(self):
        return self._pattern()[1]

    def tocsr(self):
        import scipy.sparse as sps

        indptr, indices = self._pattern()
        return sps.csr_matrix((self.data.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()), shape=self.shape)

    def toarray(self):
        return self.tocsr().toarray()

    def __repr__(self):
        return f"DeviceCSR({self.shape[0]}x{self.shape[1]}, nnz={self.nnz}, {self.grid})"
