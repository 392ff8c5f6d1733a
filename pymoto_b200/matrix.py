"""DeviceCSR: the assembled system matrix as it lives in HBM.

It is the reference's CSR matrix (``csr_matrix((data, indices, indptr))``, pymoto/modules/assembly.py:275) with
the ``data`` array resident on the GPU.  On a structured voxel grid ``indptr``/``indices`` are a closed form of
the grid, so they are only materialised (bit-exactly, int32 when nnz fits like scipy's downcast) when a user asks
for them: ``.indptr``, ``.indices`` or ``.tocsr()``.  The solver kernels stream ``data`` alone.

With a slab decomposition (pymoto_b200/slab.py) the object holds the rows of this rank's node planes; vectors it
multiplies are views into buffers padded by one halo plane on each side (:meth:`new_vec`), refreshed from the
neighbours by :meth:`exchange` before every operator application.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from . import device as dv


def make_grid(nx, ny, nz, ndof, kz0=0, nzl=None):
    return _lib.Grid(int(nx), int(ny), int(nz), int(ndof), int(kz0), int(nz + 1 if nzl is None else nzl))


class ElemGenerator:
    """Matrix-free description of a finest-level operator: the ``pmb_elem_op`` handed to ``pmb_elem_spmv`` plus the tensors
    that keep its pointers alive.  The kernel layout (``variant``) lives here, per operator; nothing is process-global in
    the library.  ``gen["ke" | "s" | "mask" | "bcdiag"]`` reads the parts."""

    tuned = {}  # ndof -> layout measured by pmb_elem_autotune on the first large operator (cache of a measurement)

    def __init__(self, grid, ke, s, mask, bcdiag, bc=None):
        self.grid, self.ke, self.s, self.mask, self.bcdiag = grid, ke, s, mask, float(bcdiag)
        self.bc = bc  # GLOBAL Dirichlet dof numbers (host array) behind ``mask``, or None
        env = os.environ.get("PMB_ELEM_VARIANT")
        self.variant = int(env) if env is not None else ElemGenerator.tuned.get(grid.ndof, 0)
        self._flags = {}  # (kz0, nzl) -> brick flags of that (sub-)slab
        self._ops = {}

    def __getitem__(self, key):
        return getattr(self, key)

    def retarget(self, s):
        """New element scaling vector (same storage layout): cached descriptors are rebuilt only if the address moved."""
        if s.data_ptr() != self.s.data_ptr():
            self._ops.clear()
        self.s = s

    def op(self, sub=None, k_rel=0):
        """``pmb_elem_op`` for the whole slab, or for the sub-slab ``sub`` starting ``k_rel`` planes above its first plane."""
        g = self.grid if sub is None else sub
        key = (g.kz0, g.nzl, self.variant)
        o = self._ops.get(key)
        if o is None:
            lay, plane = g.nx * g.ny, (g.nx + 1) * (g.ny + 1) * g.ndof
            mask_ptr = None if self.mask is None else self.mask.data_ptr() + k_rel * plane
            flags = None
            fv = 6 if (self.variant in (6, 7) and g.ndof == 3) else (4 if self.variant in (4, 5) else None)
            if self.variant >= 8:
                fv = self.variant
            if mask_ptr is not None and g.nz > 0 and g.ndof != 2 and fv is not None:
                flags = self._flags.get(key[:2] + (fv,))
                if flags is None:
                    flags = self._flags[key[:2] + (fv,)] = dv.empty(_lib.query("pmb_elem_brickflags_bytes", g, fv), torch.uint8)
                    _lib.call("pmb_elem_brickflags", g, fv, mask_ptr, flags.data_ptr(), dv.stream())
            o = self._ops[key] = _lib.ElemOp(self.ke.ctypes.data, self.s.data_ptr() + 8 * k_rel * lay, mask_ptr, self.bcdiag,
                                             None if flags is None else flags.data_ptr(), int(self.variant))
        return o


class DeviceCSR:
    # Distributed operator applications without fused dot products are split into interior planes (launched while the
    # halo planes are in flight) and the two boundary planes (launched once they arrived).
    overlap_halo = os.environ.get("PMB_OVERLAP_HALO", "1") != "0"
    overlap_min_rows = 3_000_000  # below this the three sub-launches cost more host time than the exchange they hide

    # Apply finest-level operators matrix-free (from the element scaling vector, pmb_elem.cu) when the assembly module
    # attached its generator; the assembled values stay the source of truth for everything else.  Set to False to
    # stream the CSR values on every level.
    matrix_free = True
    # The 3-D matrix-free kernel exists in several layouts (pmb_elem.cu: brick / z-marching columns, bit-identical in y): on
    # operators of at least ``autotune_min_rows`` rows the first assembly times them once on its own operands and keeps
    # the fastest for the process (PMB_ELEM_VARIANT=<n> pins one instead; PMB_ELEM_AUTOTUNE=0 keeps variant 0).
    autotune_min_rows = 1_000_000
    # The FP64 tensor-core layouts (3, 6) accumulate in a different order: their y equals the DFMA layouts' to rounding
    # (1e-16 relative per term), not bit for bit.  Everything the reference pins (residual 1e-8, compliance 1e-6) is
    # far above that, so the autotune may select them; PMB_ELEM_BITEXACT=1 restricts it to the bit-identical layouts.
    allow_rounding_layouts = os.environ.get("PMB_ELEM_BITEXACT", "0") != "1"
    elem_timings_ms = {}  # filled by the autotune pass: ndof -> [ms per launch of every layout]
    # Coarse-level operators (level >= 1, whole 3-D grids of at least ``symmetric_min_nodes`` nodes) are ALSO kept in the
    # symmetric half-stencil layout of pmb_symstore.cu (14 instead of 27 blocks per node) and swept from it; the stencil-CSR
    # values stay the source of truth (Galerkin products, diagonal, densify).  PMB_SYMMETRIC_STORAGE=0 switches it off.
    symmetric_storage = os.environ.get("PMB_SYMMETRIC_STORAGE", "0") == "1"
    symmetric_min_nodes = 100_000
    symmetric_tol = 1e-10  # max |A_ij - A_ji^T| / max |A_ij| above which the operator is not treated as symmetric

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
        self._diag_buf = self._nnz_off_buf = None
        self._entry_offsets = {}
        self.generator = None  # ElemGenerator (set by the assembly module) or None
        self._sym = self._sym_stats = None  # symmetric half-stencil copy of the values (pack_symmetric) + its asymmetry stats
        self._sym_valid = False

    # ---- values
    @property
    def data(self):
        return self._buf[: self.nnz]

    def invalidate(self):
        """Call after the values changed in place."""
        self._diag = self._nnz_off = None
        self._sym_valid = False

    def pack_symmetric(self):
        """(Re)build the symmetric half-stencil copy of the current values if this operator qualifies (coarse level, whole
        3-D grid, large enough, symmetric to ``symmetric_tol``); the sweeps then read 14 instead of 27 blocks per node.
        Reads two 8-byte statistics back (one host synchronisation per update)."""
        g = self.grid
        self._sym_valid = False
        if (not DeviceCSR.symmetric_storage or self.level < 1 or self.comm is not None or g.nz == 0 or g.kz0 != 0 or g.nzl != g.nz + 1
                or self.n // g.ndof < DeviceCSR.symmetric_min_nodes or torch.cuda.is_current_stream_capturing()):
            return False
        if self._sym is None:
            self._sym = dv.empty(_lib.query("pmb_sym_doubles", g))
            self._sym_stats = dv.empty(2, torch.int64)
        _lib.call("pmb_sym_pack", g, dv.ptr(self._buf), dv.ptr(self._sym), dv.ptr(self._sym_stats), dv.stream())
        diff, amax = self._sym_stats.cpu().numpy().view(np.float64)
        self.asymmetry = float(diff / amax) if amax > 0 else 0.0
        self._sym_valid = bool(self.asymmetry <= DeviceCSR.symmetric_tol)
        return self._sym_valid

    # ---- vectors this operator can be applied to (halo-padded), halo refresh, global dot products
    def new_vec(self, zero=False):
        """Owned entries of a fresh vector whose storage carries one halo node plane on each side (+ one node row of slack at
        the end).  The pads are always zeroed: the bulk-copy / TMA staged matrix-free layouts read them (never using the
        values, but they must be finite), and the first pad plane starts 16-byte aligned, which the tensor-map layout needs."""
        row = (self.grid.nx + 1) * self.grid.ndof
        base = dv.empty(self.n + 2 * self.plane + row + 2)
        if zero:
            base.zero_()
        else:
            base[: self.plane] = 0.0
            base[self.plane + self.n:] = 0.0
        base._pmb_padded = self.plane
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

    def _padded(self, x):
        """``x`` if it is a view made by :meth:`new_vec` (plane-padded, aligned, finite pads), else a padded copy."""
        base = x._base
        if base is not None and getattr(base, "_pmb_padded", 0) == self.plane and x.storage_offset() == self.plane and x.numel() == self.n:
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
    def rowstats_buffers(self):
        if self._diag_buf is None:  # persistent storage: addresses stay fixed across updates (CUDA-graph replays rely on it)
            self._diag_buf = dv.empty(self.n)
            self._nnz_off_buf = dv.empty(self.n, torch.int32)
        return self._diag_buf, self._nnz_off_buf

    def rowstats(self):
        if self._diag is None:
            self._diag, self._nnz_off = self.rowstats_buffers()
            _lib.call("pmb_rowstats", self.grid, dv.ptr(self._buf), dv.ptr(self._diag), dv.ptr(self._nnz_off), dv.stream())
        return self._diag, self._nnz_off

    def diagonal_device(self):
        return self.rowstats()[0]

    def diagonal(self):
        return self.diagonal_device().cpu().numpy()

    def autotune_matrix_free(self):
        """Measure the matrix-free kernel layouts on this operator (once per process and dofs per node) and keep the fastest
        of the bit-identical ones; no-op when not applicable (2-D, ndof 2, small operators, a pinned PMB_ELEM_VARIANT)."""
        g, gen = self.grid, self.generator
        if gen is None or g.ndof in ElemGenerator.tuned:
            return
        if (not DeviceCSR.matrix_free or g.nz == 0 or g.ndof == 2 or self.n < DeviceCSR.autotune_min_rows
                or "PMB_ELEM_VARIANT" in os.environ or os.environ.get("PMB_ELEM_AUTOTUNE", "1") == "0"
                or torch.cuda.is_current_stream_capturing()):
            return
        x, b, y = self.new_vec(zero=True), self.new_vec(zero=True), self.new_vec(zero=True)
        x.copy_(torch.rand(self.n, dtype=torch.float64, device=x.device))
        ms = (C.c_double * _lib.query("pmb_elem_num_variants"))()
        best = C.c_int(0)
        scratch = None if gen.mask is None else dv.empty(_lib.query("pmb_elem_autotune_flag_bytes", g), torch.uint8)
        _lib.call("pmb_elem_autotune", g, C.byref(gen.op()), x.data_ptr(), b.data_ptr(), self.diagonal_device().data_ptr(),
                  y.data_ptr(), dv.ptr(scratch), 1 if DeviceCSR.allow_rounding_layouts else 0, C.addressof(ms), C.byref(best),
                  dv.stream())
        DeviceCSR.elem_timings_ms[g.ndof] = [float(v) for v in ms]
        ElemGenerator.tuned[g.ndof] = gen.variant = int(best.value)

    # ---- products
    def _launch(self, gen, grid, mode, data_ptr, k_rel, x_ptr, b_ptr, diag_ptr, w, y_ptr, dotv_ptr, dot_ptr, ws_ptr):
        if gen is None and self._sym_valid and grid is self.grid and DeviceCSR.symmetric_storage:
            _lib.call("pmb_sym_spmv", grid, mode, dv.ptr(self._sym), x_ptr, b_ptr, diag_ptr, float(w), y_ptr, dotv_ptr, dot_ptr, ws_ptr,
                      dv.stream())
        elif gen is None:
            _lib.call("pmb_spmv", grid, mode, data_ptr, x_ptr, b_ptr, diag_ptr, float(w), y_ptr, dotv_ptr, dot_ptr, ws_ptr,
                      dv.stream())
        else:
            op = gen.op() if grid is self.grid else gen.op(grid, k_rel)
            _lib.call("pmb_elem_spmv", grid, mode, C.byref(op), x_ptr, b_ptr, diag_ptr, float(w), y_ptr, dotv_ptr, dot_ptr, ws_ptr,
                      dv.stream())

    def apply(self, mode, x, y, b=None, diag=None, w=0.0, dotv=None, dot_out=None):
        """Raw kernel call on device tensors: y = A x | b - A x | x + w (b - A x)/diag, optional fused dots."""
        gen = self.generator if DeviceCSR.matrix_free else None
        g = self.grid
        ws = None
        if dot_out is not None:
            ws = dv.workspace().spmv_ws(_lib.query("pmb_spmv_ws_doubles" if gen is None else "pmb_elem_ws_doubles", g))

        def P(t):
            return None if t is None else t.data_ptr()

        if (self.comm is None or dot_out is not None or not DeviceCSR.overlap_halo or g.nzl < 3 or self.n < DeviceCSR.overlap_min_rows
                or self.comm.fast):  # mailbox exchanges are cheap enough to stay on the compute stream
            if self.comm is not None:
                self.exchange(x)
            if gen is not None:
                x = self._padded(x)  # the bulk-copy layouts read whole 16-byte granules around every staged row
            self._launch(gen, g, mode, P(self._buf), 0, P(x), P(b), P(diag), w, P(y), P(dotv), P(dot_out), P(ws))
            if dot_out is not None and self.comm is not None:
                self.comm.allreduce_(dot_out)
            return y
        # ---- distributed, overlapped: post the halo exchange, compute the interior planes, then the two boundary planes
        base = x._base if x._base is not None else x
        off = x.storage_offset()
        if x._base is None or off < self.plane or base.numel() < off + self.n + self.plane:
            raise _lib.PmbError("distributed operator input must come from DeviceCSR.new_vec() (halo-padded storage)")
        reqs = self.comm.exchange_start(base, off, self.n, self.plane)
        def sub(k_rel, nplanes):
            """Launch on owned planes [k_rel, k_rel + nplanes) (relative to kz0): every pointer moves with the sub-slab."""
            sg = make_grid(g.nx, g.ny, g.nz, g.ndof, g.kz0 + k_rel, nplanes)
            dofs = k_rel * self.plane
            first_entry = self._entry_offsets.get(k_rel)
            if first_entry is None:
                first_entry = 0 if k_rel == 0 else _lib.query("pmb_nnz", make_grid(g.nx, g.ny, g.nz, g.ndof, g.kz0, k_rel))
                self._entry_offsets[k_rel] = first_entry
            sh = lambda p, n, sz: None if p is None else p + n * sz  # noqa: E731
            self._launch(gen, sg, mode, sh(P(self._buf), first_entry, 8), k_rel,
                         sh(P(x), dofs, 8), sh(P(b), dofs, 8), sh(P(diag), dofs, 8), w, sh(P(y), dofs, 8), None, None, None)

        sub(1, g.nzl - 2)
        self.comm.exchange_finish(reqs)
        sub(0, 1)
        sub(g.nzl - 1, 1)
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
