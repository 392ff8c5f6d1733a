"""Linear solvers of the hot path on the GPU: LDAS wrapper, PCG, geometric multigrid, damped Jacobi, and the
dense coarsest-level solver.

Mirrors the ``LinearSolver`` plug-in API of the reference (pymoto/solvers/solvers.py:6-50: ``update(A)``,
``solve(rhs, x0=None, trans="N")``) and the classes

  LDAWrapper           pymoto/solvers/solvers.py:99-306
  DampedJacobi         pymoto/solvers/iterative.py:21-47
  GeometricMultigrid   pymoto/solvers/iterative.py:124-256
  CG                   pymoto/solvers/iterative.py:295-403

with the same constructor signatures and the same iteration (so iteration counts match the reference).  All
vectors inside the solvers are CUDA tensors; ``solve`` accepts numpy arrays or CUDA tensors and returns the same
kind.  The matrices are :class:`DeviceCSR` (real, symmetric), so ``trans`` in {"N", "T", "H"} all solve the same
system; anything else raises ``TypeError`` like the reference.
"""
import os
import time
import warnings

import numpy as np
import torch

from . import _lib
from . import device as dv
from .domain import VoxelDomain, grid_dims
from .matrix import DeviceCSR, make_grid


def _check_trans(trans):
    if trans not in ("N", "T", "H"):
        raise TypeError("Only N, T, or H transposition is possible")


def _check_matrix(A, who):
    if not isinstance(A, DeviceCSR):
        raise TypeError(f"{who} works on a pymoto_b200.DeviceCSR (got {type(A).__name__}); there is no CPU fallback")


class LinearSolver:
    """Base class (pymoto/solvers/solvers.py:6-50)."""

    defined = True
    _err_msg = ""

    def __init__(self, A=None):
        if A is not None:
            self.update(A)

    def update(self, A):
        raise NotImplementedError(f"Solver not implemented {self._err_msg}")

    def solve(self, rhs, x0=None, trans="N"):
        raise NotImplementedError(f"Solver not implemented {self._err_msg}")

    @staticmethod
    def residual(A, x, b, trans="N"):
        """Relative residual |A x - b| / |b| (|b| = 0 -> absolute), pymoto/solvers/solvers.py:52-85."""
        _check_trans(trans)
        xd, bd = dv.to_device(x).reshape(-1), dv.to_device(b).reshape(-1)
        assert xd.shape == bd.shape
        r = dv.empty(bd.numel())
        A.apply(_lib.RESIDUAL, A.operand(xd), r, b=bd)
        d = A.dots([(r, r), (bd, bd)]).cpu().numpy()
        bnorm = np.sqrt(d[1]) if d[1] != 0 else 1.0
        return float(np.sqrt(d[0]) / bnorm)


class Preconditioner(LinearSolver):
    """Identity preconditioner (pymoto/solvers/iterative.py:11-18)."""

    def update(self, A):
        pass

    def solve(self, rhs, x0=None, trans="N"):
        return rhs.clone() if isinstance(rhs, torch.Tensor) else rhs.copy()


class DampedJacobi(Preconditioner):
    r"""``M = D / w`` (pymoto/solvers/iterative.py:21-47)."""

    def __init__(self, A=None, w=1.0):
        assert 0 < w <= 1, "w must be between 0 and 1"
        self.w = w
        self.D = None
        super().__init__(A)

    def update(self, A):
        _check_matrix(A, "DampedJacobi")
        self.D = A.diagonal_device()

    def solve(self, rhs, x0=None, trans="N"):
        _check_trans(trans)
        r = dv.to_device(rhs).reshape(-1)
        u = dv.empty(r.numel())
        _lib.call("pmb_smooth0", r.numel(), float(self.w), dv.ptr(r), dv.ptr(self.D), dv.ptr(u), dv.stream())
        return dv.like_input(u, rhs)


class SolverDenseInverse(LinearSolver):
    """Coarsest-level direct solver: explicit FP64 inverse of the (small, SPD) coarse operator, applied as a GEMV.

    Stands where the reference's ``auto_determine_solver`` puts a sparse LU (pymoto/solvers/iterative.py:174-176,
    pymoto/solvers/sparse.py:533-550); 675 x 675 for the 3-D elasticity configurations.
    """

    max_size = 8192

    def __init__(self, A=None):
        self.n = 0
        self.inv = None
        super().__init__(A)

    def update(self, A):
        _check_matrix(A, "SolverDenseInverse")
        n = A.shape[0]
        if n > self.max_size:
            raise ValueError(f"Coarsest multigrid level has {n} dofs, too many for the dense direct solve "
                             f"(max {self.max_size}); add GeometricMultigrid levels")
        if self.inv is None or self.n != n:
            self.n = n
            self.inv = dv.empty(n * n)
            self._scratch = dv.empty(_lib.query("pmb_dense_invert_ws_doubles", n))
            self._info = dv.zeros(1, torch.int32)
            self._out = dv.empty(n)
        st = dv.stream()
        _lib.call("pmb_densify", A.grid, dv.ptr(A._buf), dv.ptr(self.inv), st)
        _lib.call("pmb_dense_invert", n, dv.ptr(self.inv), dv.ptr(self._scratch), dv.ptr(self._info), st)
        self._checked = False

    def _check_info(self):
        if not self._checked:
            info = int(self._info.item())
            if info != 0:
                raise np.linalg.LinAlgError(f"coarsest-level operator is not positive definite (pivot {info - 1})")
            self._checked = True

    def solve(self, rhs, x0=None, trans="N"):
        _check_trans(trans)
        if not self._checked and not torch.cuda.is_current_stream_capturing():
            self._check_info()  # first solve after an update: a non-positive pivot must not go unnoticed (the reference pivots)
        r = dv.to_device(rhs).reshape(-1)
        out = self._out if dv.is_device(rhs) else dv.empty(self.n)
        _lib.call("pmb_dense_gemv", self.n, dv.ptr(self.inv), dv.ptr(r), dv.ptr(out), dv.stream())
        return dv.like_input(out, rhs)


class GeometricMultigrid(Preconditioner):
    """Geometric multigrid preconditioner, one V-cycle per ``solve`` (pymoto/solvers/iterative.py:124-256).

    Trilinear prolongation (weights 1, 1/2, 1/4, 1/8, no Dirichlet awareness), restriction = its transpose,
    Galerkin coarse operator rebuilt on every ``update``, ``smooth_steps`` damped-Jacobi sweeps before and after
    the coarse correction (the first pre-sweep starts from zero: ``u = w r / D``).
    """

    _available_cycles = ["v", "w"]
    # Replay the whole V-cycle (all levels, ~70 kernel launches) as ONE CUDA graph when it is used as a preconditioner on
    # a single GPU.  Every buffer the cycle touches is persistent, so the graph is captured once per hierarchy and only
    # re-captured if an address changes.  PMB_CUDA_GRAPH=0 disables it.
    use_cuda_graph = os.environ.get("PMB_CUDA_GRAPH", "1") != "0"
    # Build the level-1 operator straight from the element densities (pmb_galerkin_direct, pymoto_b200/coarse.py) when the
    # fine matrix comes from an assembly module (3-D): the fine values are not read and no intermediate is written.
    # PMB_GALERKIN_DIRECT=0 keeps the generic two-pass product R^T (A R) on every level.
    direct_level1 = os.environ.get("PMB_GALERKIN_DIRECT", "1") != "0"
    # Capture the V-cycle of a slab-decomposed hierarchy as well (needs the symmetric-memory mailboxes: the halo copies, the
    # device-side barriers and the peer-store gather of the first replicated level are ordinary stream work).
    graph_on_slabs = os.environ.get("PMB_GRAPH_SLABS", "1") != "0"

    def __init__(self, domain, A=None, cycle: str = "V", inner_level: LinearSolver = None, smoother: LinearSolver = None,
                 smooth_steps: int = 5):
        nx, ny, nz = grid_dims(domain)
        assert nx % 2 == 0 and ny % 2 == 0 and nz % 2 == 0, f"Domain sizes {nx, ny, nz} must be divisible by 2"
        self.domain = domain
        self.A = None
        assert cycle.lower() in self._available_cycles, f"Cycle ({cycle}) is not available. Options are {self._available_cycles}"
        self.cycle = cycle
        self.inner_level = inner_level
        self.smoother = DampedJacobi(w=0.5) if smoother is None else smoother
        if not isinstance(self.smoother, DampedJacobi):
            raise TypeError("pymoto_b200.GeometricMultigrid smooths with DampedJacobi (fused sweep kernel) only")
        self.smooth_steps = smooth_steps
        self.sub_domain = VoxelDomain(nx // 2, ny // 2, nz // 2, domain.unitx * 2, domain.unity * 2, domain.unitz * 2)
        self.Ac = None
        self._buf = None
        self._fine_key = None
        self._replicate = False
        self._graph = None
        self._graph_sig = None
        self._graph_failed = False
        self._rc_sym = None
        self._eager_calls = 0
        super().__init__(A)

    def update(self, A):
        _check_matrix(A, "GeometricMultigrid")
        g = A.grid
        nx, ny, nz = grid_dims(self.domain)
        if (g.nx, g.ny, g.nz) != (nx, ny, nz):
            raise ValueError(f"Matrix grid {(g.nx, g.ny, g.nz)} does not match the multigrid domain {(nx, ny, nz)}")
        self.A = A
        self.smoother.update(A)
        if self.Ac is None or self.Ac.grid.ndof != g.ndof or self._fine_key != (g.kz0, g.nzl, A.comm is not None):
            self._fine_key = (g.kz0, g.nzl, A.comm is not None)
            self._setup_coarse(A)
        gcl = self._gc_local
        st = dv.stream()
        gen = A.generator
        if GeometricMultigrid.direct_level1 and A.level == 0 and gen is not None and g.nz > 0:
            t = self._direct_tables(gen, gcl)
            _lib.call("pmb_galerkin_direct", g, gcl, dv.ptr(t["Gtab"]), dv.ptr(t["cidx"]), dv.ptr(t["child_ids"]), gen.s.data_ptr(),
                      dv.ptr(self._Ac_local._buf), st)
            if t["bc_idx"] is not None:
                _lib.call("pmb_scatter_add", t["bc_idx"].numel(), dv.ptr(t["bc_idx"]), dv.ptr(t["bc_val"]), dv.ptr(self._Ac_local._buf), st)
        else:
            # two streaming passes; between them the lower halo plane of the column-collapsed intermediate is fetched
            # from the rank below (its top owned fine plane contributes to my first coarse plane)
            nwork = _lib.query("pmb_galerkin_ws_doubles", g)
            bplane = nwork // g.nzl  # one node plane of the intermediate
            work = dv.workspace().galerkin_ws(nwork + bplane)
            _lib.call("pmb_galerkin_cols", g, gcl, dv.ptr(A._buf), work.data_ptr() + 8 * bplane, st)
            if A.comm is not None:
                A.comm.exchange(work, bplane, nwork, bplane, lower=True, upper=False)
            _lib.call("pmb_galerkin_rows", g, gcl, work.data_ptr() + 8 * bplane, dv.ptr(self._Ac_local._buf), st)
        if self._replicate:  # first replicated level: every rank gets the whole coarse operator
            A.comm.gather_full(self._Ac_local.data, self.Ac.data, self._coarse_entry_offset)
        self.Ac.invalidate()
        self.Ac.pack_symmetric()  # levels >= 1 are swept from the symmetric half-stencil copy where it applies
        if self.inner_level is None:
            self.inner_level = SolverDenseInverse()
        self.inner_level.update(self.Ac)

    def _direct_tables(self, gen, gcl):
        """Device copies of the constant tables of the direct level-1 build (once per element matrix / bc set / slab)."""
        from . import coarse

        key = (id(gen), gen.ke.tobytes(), gen.bcdiag, gcl.kz0, gcl.nzl)
        if getattr(self, "_direct_key", None) != key:
            g = gen.grid
            h = coarse.build_tables(gen.ke, g.ndof, (g.nx, g.ny, g.nz), gen.bc if gen.mask is not None else None, gen.bcdiag,
                                    gcl.kz0, gcl.kz0 + gcl.nzl)
            up = lambda a, dt: None if a is None or a.size == 0 else dv.to_device(a, dt)  # noqa: E731
            self._direct = dict(Gtab=dv.to_device(h["Gtab"].ravel()), cidx=up(h["cidx"], torch.int32),
                                child_ids=None if h["child_ids"] is None else dv.to_device(np.ascontiguousarray(h["child_ids"]).view(np.int16).ravel(), torch.int16),
                                bc_idx=up(h["bc_idx"], torch.int64), bc_val=up(h["bc_val"], torch.float64))
            self._direct_key = key
        return self._direct

    def _setup_coarse(self, A):
        """Coarse operator storage and level buffers (once per problem)."""
        from . import slab

        g = A.grid
        nx, ny, nz = g.nx, g.ny, g.nz
        n = A.shape[0]
        if A.comm is None:  # this level lives on one GPU (single-GPU run, or a replicated level)
            self._gc_local = make_grid(nx // 2, ny // 2, nz // 2, g.ndof)
            self.Ac = self._Ac_local = DeviceCSR(self._gc_local, level=A.level + 1)
            self._replicate = False
        else:
            part = slab.context().part
            k0c, k1c = part.slab_planes(A.level + 1)
            self._gc_local = make_grid(nx // 2, ny // 2, nz // 2, g.ndof, k0c, k1c - k0c)
            self._replicate = not part.is_distributed(A.level + 1)
            if self._replicate:
                self._Ac_local = DeviceCSR(self._gc_local)
                self.Ac = DeviceCSR(make_grid(nx // 2, ny // 2, nz // 2, g.ndof), level=A.level + 1)
                self._coarse_entry_offset = 0 if k0c == 0 else _lib.query(
                    "pmb_nnz", make_grid(nx // 2, ny // 2, nz // 2, g.ndof, 0, k0c))
                self._coarse_row_offset = k0c * self.Ac.plane
                self._rc_sym = None
                if A.comm.fast and os.environ.get("PMB_SYMM_GATHER", "1") != "0":
                    try:  # replicated right-hand side assembled by peer stores (no collective, capturable in a CUDA graph)
                        self._rc_full, peers, hdl = A.comm.symmetric_vector(self.Ac.shape[0])
                        self._rc_sym = (peers, hdl)
                    except Exception as e:
                        warnings.warn(f"symmetric-memory gather unavailable ({e}); using the NCCL all-reduce")
                if self._rc_sym is None:
                    self._rc_full = dv.empty(self.Ac.shape[0])
            else:
                self.Ac = self._Ac_local = DeviceCSR(self._gc_local, comm=A.comm, level=A.level + 1)
        nc_local = _lib.query("pmb_nrows", self._gc_local)
        self._buf = dict(u=A.new_vec(), u2=A.new_vec(), t=A.new_vec(), rc=dv.empty(nc_local))

    def _signature(self):
        """Addresses of everything a captured V-cycle reads or writes, down the whole hierarchy."""
        sig, lvl = [], self
        while isinstance(lvl, GeometricMultigrid):
            if lvl.A is None or lvl._buf is None:
                return None
            if lvl.A.comm is not None and not (lvl.A.comm.fast and GeometricMultigrid.graph_on_slabs
                                               and (not lvl._replicate or lvl._rc_sym is not None)):
                return None  # slab levels are capturable only with mailbox halos and the peer-store gather
            gen = lvl.A.generator if DeviceCSR.matrix_free else None
            sig += [lvl.A._buf.data_ptr(), lvl.smoother.D.data_ptr(), float(lvl.smoother.w), lvl.smooth_steps,
                    None if gen is None else (gen["s"].data_ptr(), None if gen["mask"] is None else gen["mask"].data_ptr(),
                                              gen["bcdiag"], gen["ke"].ctypes.data, gen["ke"].tobytes(), gen.variant)]
            sig += [lvl._buf[k].data_ptr() for k in ("u", "u2", "t", "rc")]
            sig += [lvl.A.grid.kz0, lvl.A.grid.nzl, lvl._replicate]
            sig += [lvl.A._sym.data_ptr() if (lvl.A._sym_valid and DeviceCSR.symmetric_storage) else None]
            lvl = lvl.inner_level
        if not isinstance(lvl, SolverDenseInverse) or lvl.inv is None:
            return None
        return tuple(sig + [lvl.inv.data_ptr(), lvl._out.data_ptr(), lvl.n])

    def _coarsest(self):
        lvl = self
        while isinstance(lvl, GeometricMultigrid):
            lvl = lvl.inner_level
        return lvl if isinstance(lvl, SolverDenseInverse) else None

    def solve(self, rhs, x0=None, trans="N"):
        _check_trans(trans)
        if self.A is not None and self.A.level == 0 and not torch.cuda.is_current_stream_capturing():
            inner = self._coarsest()  # graph replays never pass through SolverDenseInverse.solve: check the pivots here
            if inner is not None and inner.inv is not None and not inner._checked:
                inner._check_info()
        if (GeometricMultigrid.use_cuda_graph and x0 is None and dv.is_device(rhs) and self.A is not None
                and self.A.level == 0 and rhs.numel() == self.A.shape[0] and not self._graph_failed
                and not torch.cuda.is_current_stream_capturing()):
            sig = self._signature()
            if sig is not None:
                if self._graph is not None and sig != self._graph_sig:
                    self._graph, self._eager_calls = None, 0  # re-measure the launch statistics with one eager pass, then re-capture
                if self._graph is None and self._eager_calls >= 1:  # one eager pass first (lazy kernel attributes etc.)
                    comm = self.A.comm
                    self._graph_in = dv.empty(self.A.shape[0])
                    self._graph_in.copy_(rhs.reshape(-1))
                    torch.cuda.synchronize()
                    try:
                        g = torch.cuda.CUDAGraph()
                        if comm is not None:  # every rank captures the same V-cycle, mailbox exchanges and barriers included
                            comm.barrier()
                            comm.begin_capture()
                        with torch.cuda.graph(g):
                            self._graph_out = self._solve_eager(self._graph_in, None)
                            if comm is not None:
                                comm.end_capture()
                        self._graph, self._graph_sig = g, self._signature()
                    except Exception as e:  # e.g. a symmetric-memory barrier that cannot be captured on this build
                        if comm is None:
                            raise
                        warnings.warn(f"slab V-cycle could not be captured into a CUDA graph ({e}); running it eagerly")
                        self._graph, self._graph_failed = None, True
                        torch.cuda.synchronize()
                if self._graph is not None:
                    self._graph_in.copy_(rhs.reshape(-1))
                    self._graph.replay()
                    _lib.launch_count += self._graph_launches  # the replay launches the same kernels as the eager pass
                    for key, cnt in self._graph_stats.items():
                        _lib.call_stats[key] = _lib.call_stats.get(key, 0) + cnt
                    if self.A.comm is not None:
                        self.A.comm.exchanges += self._graph_comm[0]
                        self.A.comm.allreduces += self._graph_comm[1]
                    return self._graph_out
        self._eager_calls += 1
        n0, s0 = _lib.launch_count, dict(_lib.call_stats)
        c0 = (self.A.comm.exchanges, self.A.comm.allreduces) if self.A is not None and self.A.comm is not None else (0, 0)
        out = self._solve_eager(rhs, x0)
        if self.A is not None and self.A.comm is not None:
            self._graph_comm = (self.A.comm.exchanges - c0[0], self.A.comm.allreduces - c0[1])
        self._graph_launches = _lib.launch_count - n0
        self._graph_stats = {k: c - s0.get(k, 0) for k, c in _lib.call_stats.items() if c - s0.get(k, 0) > 0}
        return out

    def _solve_eager(self, rhs, x0):
        b = dv.to_device(rhs).reshape(-1)
        A, D, w = self.A, self.smoother.D, float(self.smoother.w)
        n = A.shape[0]
        st = dv.stream()
        u, u2, t, rc = self._buf["u"], self._buf["u2"], self._buf["t"], self._buf["rc"]
        # pre-smoothing
        if x0 is None:
            _lib.call("pmb_smooth0", n, w, dv.ptr(b), dv.ptr(D), dv.ptr(u), st)
        else:
            A.apply(_lib.JACOBI, A.operand(dv.to_device(x0).reshape(-1)), u, b=b, diag=D, w=w)
        for _ in range(self.smooth_steps - 1):
            A.apply(_lib.JACOBI, u, u2, b=b, diag=D, w=w)
            u, u2 = u2, u
        # coarse-grid correction
        A.apply(_lib.RESIDUAL, u, t, b=b)
        A.exchange(t, lower=True, upper=False)  # restriction of coarse plane K reads fine planes 2K-1 .. 2K+1
        _lib.call("pmb_restrict", A.grid, self._gc_local, dv.ptr(t), dv.ptr(rc), st)
        if self._replicate:
            if self._rc_sym is not None:
                A.comm.bcast_part(rc, self._rc_sym[0], self._rc_sym[1], self._coarse_row_offset)
            else:
                A.comm.gather_full(rc, self._rc_full, self._coarse_row_offset)
            uc_full = self.inner_level.solve(self._rc_full)
            uc_ptr = uc_full.data_ptr() + 8 * self._coarse_row_offset  # the slab view includes its upper halo plane
        else:
            uc = self.inner_level.solve(rc)
            if self.Ac.comm is not None:
                uc = self.Ac.operand(uc)
                self.Ac.exchange(uc, lower=False, upper=True)  # odd fine planes interpolate from coarse plane K+1
            uc_ptr = uc.data_ptr()
        _lib.call("pmb_prolong_add", A.grid, self._gc_local, uc_ptr, dv.ptr(u), st)
        # post-smoothing
        for _ in range(self.smooth_steps):
            A.apply(_lib.JACOBI, u, u2, b=b, diag=D, w=w)
            u, u2 = u2, u
        self._buf["u"], self._buf["u2"] = u, u2  # the result lives in an internal buffer until the next solve()
        return dv.like_input(u, rhs)


def auto_multigrid(domain, min_size=8):
    """The level chain of examples/topology_optimization/ex_compliance_multigrid.py:107-121: coarsen while every
    coarse dimension stays even and >= ``min_size``.  Returns the list of levels, finest first."""
    mgs = [GeometricMultigrid(domain)]
    while True:
        size = mgs[-1].sub_domain.size
        if any(n % 2 != 0 for n in size) or any(size < min_size):
            break
        mgs.append(GeometricMultigrid(mgs[-1].sub_domain))
        mgs[-2].inner_level = mgs[-1]
    return mgs


class CG(LinearSolver):
    """Preconditioned conjugate gradients, the reference's variant (pymoto/solvers/iterative.py:295-403):
    ``p0 = z/|z|``, ``alpha = (p.q)^-1 p.r``, ``beta = -(p.q)^-1 q.z``, explicit residual every ``restart``
    iterations (including the first), stop on ``|r|/|b| <= tol``.  One right-hand side at a time."""

    def __init__(self, A=None, preconditioner: Preconditioner = None, tol: float = 1e-7, maxit: int = 10000,
                 restart: int = 50, verbosity: int = 0):
        self.preconditioner = Preconditioner() if preconditioner is None else preconditioner
        self.A = None
        self.tol = tol
        self.maxit = maxit
        self.restart = restart
        self.verbosity = verbosity
        self.iterations = 0  # products q = A p of the last solve (the reference prints this minus one)
        self._c_plan, self._c_plan_key, self._c_bufs = None, None, None  # C driver: pmb_pcg_plan + the vectors bound to it
        self.last_residual = None
        super().__init__(A)

    def update(self, A):
        _check_matrix(A, "CG")
        tstart = time.perf_counter()
        self.A = A
        self.preconditioner.update(A)
        if self.verbosity >= 1:
            torch.cuda.synchronize()
            print(f"CG Preconditioner set up in {np.round(time.perf_counter() - tstart, 3)}s")

    def solve(self, rhs, x0=None, trans="N"):
        _check_trans(trans)
        bd = dv.to_device(rhs)
        if bd.ndim == 2:
            cols = []
            for i in range(bd.shape[1]):
                x0i = None if x0 is None else dv.to_device(x0)[:, i].contiguous()
                cols.append(self._solve1(bd[:, i].contiguous(), x0i).clone())
            return dv.like_input(torch.stack(cols, dim=1), rhs)
        return dv.like_input(self._solve1(bd.reshape(-1), None if x0 is None else dv.to_device(x0).reshape(-1)), rhs)

    # Drive the whole solve from C (pmb_pcg_plan_solve: same launches in the same order, bit-identical iterates, one host
    # poll per iteration, every non-restart CG iteration replayed as ONE CUDA graph) when the preconditioner is a single-GPU
    # geometric multigrid hierarchy.  PMB_C_PCG=1 / CG.use_c_driver = True forces it, PMB_C_PCG=0 forbids it; by default it
    # is used below ``c_driver_max_rows`` rows, where launch latency sets the pace (the Python driver replays only the
    # V-cycle as a graph and issues the rest of the iteration through ctypes, which is as fast on large grids).
    use_c_driver = os.environ.get("PMB_C_PCG", "0") == "1"

    def _mg_desc(self):
        """``pmb_mg_desc`` of the preconditioner hierarchy, or None when the C driver does not apply."""
        A, lvl = self.A, self.preconditioner
        if A is None or A.comm is not None or not isinstance(lvl, GeometricMultigrid) or lvl.A is not A:
            return None
        desc, keep, l = _lib.MgDesc(), [], 0
        while isinstance(lvl, GeometricMultigrid):
            if (l >= _lib.MAX_LEVELS or lvl.A is None or lvl.A.comm is not None or lvl._buf is None or lvl.Ac is None
                    or lvl.cycle.lower() != "v" or lvl.smoother.D is None):
                return None
            L = desc.level[l]
            L.grid = lvl.A.grid
            L.A, L.diag = lvl.A._buf.data_ptr(), lvl.smoother.D.data_ptr()
            L.u, L.u2, L.t, L.rc = (lvl._buf[k].data_ptr() for k in ("u", "u2", "t", "rc"))
            L.smooth_steps, L.w = int(lvl.smooth_steps), float(lvl.smoother.w)
            keep.append(lvl)
            coarse, lvl, l = lvl.Ac, lvl.inner_level, l + 1
        if not isinstance(lvl, SolverDenseInverse) or lvl.inv is None or lvl.n != coarse.shape[0]:
            return None
        lvl._check_info()
        desc.nlevels, desc.coarse_grid = l, coarse.grid
        desc.coarse_inv, desc.coarse_out = lvl.inv.data_ptr(), lvl._out.data_ptr()
        gen = A.generator if DeviceCSR.matrix_free else None
        if gen is not None:
            desc.gen = gen.op()
        self._desc_keep = (keep, gen)
        return desc

    def _solve1_c(self, desc, b, x0):
        import ctypes as C

        A, n = self.A, b.numel()
        tstart = time.perf_counter()
        ws = dv.workspace()
        g = A.grid
        ws_spmv = ws.spmv_ws(max(_lib.query("pmb_spmv_ws_doubles", g), _lib.query("pmb_elem_ws_doubles", g)))
        # vectors bound to the plan (persistent: the captured iteration graph holds their addresses)
        bufs = self._c_bufs
        if bufs is None or bufs["n"] != n or bufs["A"] is not A:
            bufs = self._c_bufs = dict(n=n, A=A, b=dv.empty(n), x=A.new_vec(zero=True), r=dv.empty(n), q=dv.empty(n), p=A.new_vec(),
                                       scal=dv.zeros(16))
        key = (bytes(desc), ws.red.data_ptr(), ws_spmv.data_ptr())
        if self._c_plan is None or self._c_plan_key != key:
            self._drop_plan()
            plan = C.c_void_p()
            _lib.call("pmb_pcg_plan_create", C.byref(desc), dv.ptr(bufs["b"]), dv.ptr(bufs["x"]), dv.ptr(bufs["r"]), dv.ptr(bufs["q"]),
                      dv.ptr(bufs["p"]), dv.ptr(bufs["scal"]), dv.ptr(ws.red), dv.ptr(ws_spmv), C.byref(plan))
            self._c_plan, self._c_plan_key = plan, key
        bufs["b"].copy_(b)
        if x0 is None:
            bufs["x"].zero_()
        else:
            bufs["x"].copy_(x0)
        iters, relres = C.c_int(0), C.c_double(0.0)
        launches0 = _lib.launch_count
        _lib.call("pmb_pcg_plan_solve", self._c_plan, float(self.tol), int(self.maxit), int(self.restart), C.byref(iters), C.byref(relres),
                  dv.stream())
        self.iterations, self.last_residual = int(iters.value), float(relres.value)
        # kernels launched inside the C call: per V-cycle and level 2 steps + 3 (smooth0 / residual, restrict, prolong) + the
        # coarse GEMV; per iteration the product (+ reduce), update, dots, lincomb (bench.py reports gpu_launches)
        per_cycle = sum(2 * desc.level[i].smooth_steps + 3 for i in range(desc.nlevels)) + 1
        _lib.launch_count = launches0 + 3 + self.iterations * (per_cycle + 6) + (per_cycle + 2 if self.iterations else 0)
        if self.last_residual > self.tol:
            warnings.warn(f"CG Maximum iterations ({self.maxit}) reached, with final residuals {self.last_residual}")
        elif self.verbosity >= 1:
            print(f"CG Converged in {self.iterations} iterations and {np.round(time.perf_counter() - tstart, 3)}s, "
                  f"with final (max) residual {self.last_residual}")
        x = A.new_vec()
        x.copy_(bufs["x"])
        return x

    def _drop_plan(self):
        if getattr(self, "_c_plan", None) is not None:
            _lib.call("pmb_pcg_plan_destroy", self._c_plan)
        self._c_plan, self._c_plan_key = None, None

    def __del__(self):
        try:
            self._drop_plan()
        except Exception:
            pass

    def _solve1(self, b, x0):
        A, M = self.A, self.preconditioner
        if CG.use_c_driver:
            desc = self._mg_desc()
            if desc is not None:
                return self._solve1_c(desc, b, x0)
        n = b.numel()
        tstart = time.perf_counter()
        x = A.new_vec(zero=x0 is None)
        if x0 is not None:
            x.copy_(x0)
        r, q, p = dv.empty(n), dv.empty(n), A.new_vec()
        ws = dv.workspace()
        st = dv.stream()

        A.apply(_lib.RESIDUAL, x, r, b=b)
        d = A.dots([(r, r), (b, b)]).cpu().numpy()
        bnorm = np.sqrt(d[1])
        with np.errstate(divide="ignore", invalid="ignore"):
            tval = np.sqrt(d[0]) / bnorm
        self.iterations, self.last_residual = 0, tval
        if self.verbosity >= 2:
            print(f"CG Initial (max) residual = {tval}")
        if tval <= self.tol:
            if self.verbosity >= 1:
                print(f"CG Converged in 0 iterations and {np.round(time.perf_counter() - tstart, 3)}s, "
                      f"with final (max) residual {tval}")
            return x

        z = M.solve(r, trans="N")
        zz = A.dots([(z, z)])
        dv.lincomb(p, _lib.coef(1.0, den=dv.scalar_ptr(zz), sqrt_den=True), z)  # p = z/|z|
        d3 = dv.empty(3)   # [p.q, p.r, q.r]
        rr = dv.empty(4)[:1]
        i = 0
        for i in range(self.maxit):
            A.apply(_lib.SPMV, p, q, dotv=r, dot_out=d3)
            pq, pr = dv.scalar_ptr(d3, 0), dv.scalar_ptr(d3, 1)
            if i % self.restart == 0:  # explicit residual
                _lib.call("pmb_cg_xr_update", n, dv.ptr(x), None, dv.ptr(p), None, pr, pq, None, None, st)
                A.apply(_lib.RESIDUAL, x, r, b=b)
                rr = A.dots([(r, r)])
            else:
                _lib.call("pmb_cg_xr_update", n, dv.ptr(x), dv.ptr(r), dv.ptr(p), dv.ptr(q), pr, pq, dv.ptr(rr),
                          dv.ptr(ws.red), st)
                if A.comm is not None:
                    A.comm.allreduce_(rr)
            tval = np.sqrt(float(rr[0].item())) / bnorm  # the only host sync of the iteration
            self.iterations, self.last_residual = i + 1, tval
            if self.verbosity >= 2:
                print(f"CG i = {i}, residuals = {tval}")
            if tval <= self.tol:
                break
            if not np.isfinite(tval):
                raise np.linalg.LinAlgError(f"CG residual became {tval} in iteration {i} (singular or indefinite operator / "
                                            "preconditioner, e.g. no Dirichlet condition)")
            z = M.solve(r, trans="N")
            qz = A.dots([(q, z)])
            # p = z + beta p, beta = -(q.z)/(p.q)
            dv.lincomb(p, 1.0, z, _lib.coef(-1.0, num=dv.scalar_ptr(qz), den=pq), p)

        if tval > self.tol:
            warnings.warn(f"CG Maximum iterations ({self.maxit}) reached, with final residuals {tval}")
        elif self.verbosity >= 1:
            print(f"CG Converged in {i} iterations and {np.round(time.perf_counter() - tstart, 3)}s, "
                  f"with final (max) residual {tval}")
        return x


class LDAWrapper(LinearSolver):
    """Linear-dependency-aware solver (pymoto/solvers/solvers.py:99-306) for real symmetric DeviceCSR matrices.

    Dirichlet dofs (rows whose only non-zero is the diagonal, solvers.py:88-96) are solved as ``f/diag``; the
    right-hand side is projected on the stored normalised ``(x, b = A x)`` pairs (modified Gram-Schmidt); only
    if the residual exceeds ``tol`` is the inner solver called on the remainder and the new pair stored.  The
    database is cleared on every ``update``.  For a symmetric matrix the adjoint solve (``trans="T"``) uses the
    same database, which makes the compliance adjoint cost one SpMV and no CG iteration.
    """

    def __init__(self, solver: LinearSolver, tol=1e-7, A=None, symmetric=None, hermitian=None):
        self.solver = solver
        self.tol = tol
        self.x_stored, self.b_stored = [], []
        self.A = None
        self._did_solve = False
        self._last_rtol = 0.0
        self.symmetric, self.hermitian = True, True
        self.complex = False
        super().__init__(A)

    def update(self, A, skip_inner_update: bool = False):
        _check_matrix(A, "LDAWrapper")
        self.A = A
        diag, nnz_off = A.rowstats()
        n = A.shape[0]
        self._diag = diag
        self._mask = dv.empty(n, torch.uint8)
        _lib.call("pmb_diag_mask", n, dv.ptr(diag), dv.ptr(nnz_off), dv.ptr(self._mask), dv.stream())
        self.x_stored.clear()
        self.b_stored.clear()
        if not skip_inner_update:
            self.solver.update(A)

    @property
    def diagonal_idx(self):
        return torch.nonzero(self._mask).flatten().cpu().numpy()

    def _solve1(self, rhs, x0):
        A, n, st = self.A, self.A.shape[0], dv.stream()
        sol, rhs_loc = A.new_vec(), dv.empty(n)
        _lib.call("pmb_bc_split", n, dv.ptr(self._mask), dv.ptr(rhs), dv.ptr(self._diag), dv.ptr(sol), dv.ptr(rhs_loc), st)
        # project on the database (stored vectors are zero at the Dirichlet dofs)
        for x, b in zip(self.x_stored, self.b_stored):
            d = A.dots([(rhs_loc, b), (b, b)])
            num, den = dv.scalar_ptr(d, 0), dv.scalar_ptr(d, 1)
            dv.lincomb(rhs_loc, 1.0, rhs_loc, _lib.coef(-1.0, num=num, den=den), b)
            dv.lincomb(sol, 1.0, sol, _lib.coef(1.0, num=num, den=den), x)
        self._last_rtol = self.residual(A, sol, rhs)
        self._did_solve = self._last_rtol > self.tol
        if self._did_solve:
            x0_loc = None
            if x0 is not None:
                x0_loc = dv.empty(n)
                _lib.call("pmb_mask_zero", n, dv.ptr(self._mask), dv.ptr(x0), dv.ptr(x0_loc), st)
                for x in self.x_stored:
                    d = A.dots([(x0_loc, x), (x, x)])
                    dv.lincomb(x0_loc, 1.0, x0_loc, _lib.coef(-1.0, num=dv.scalar_ptr(d, 0), den=dv.scalar_ptr(d, 1)), x)
            xnew = self.solver.solve(rhs_loc, x0=x0_loc, trans="N")
            xadd = dv.empty(n)
            _lib.call("pmb_mask_zero", n, dv.ptr(self._mask), dv.ptr(xnew), dv.ptr(xadd), st)
            dv.lincomb(sol, 1.0, sol, 1.0, xadd)
            badd = dv.empty(n)
            A.apply(_lib.SPMV, A.operand(xnew), badd)
            _lib.call("pmb_mask_zero", n, dv.ptr(self._mask), dv.ptr(badd), dv.ptr(badd), st)
            for x, b in zip(self.x_stored, self.b_stored):
                d = A.dots([(badd, b), (b, b)])
                num, den = dv.scalar_ptr(d, 0), dv.scalar_ptr(d, 1)
                dv.lincomb(badd, 1.0, badd, _lib.coef(-1.0, num=num, den=den), b)
                dv.lincomb(xadd, 1.0, xadd, _lib.coef(-1.0, num=num, den=den), x)
            bb = A.dots([(badd, badd)])
            bnrm2 = float(bb[0].item())
            if np.isfinite(bnrm2) and bnrm2 != 0:
                inv = _lib.coef(1.0, den=dv.scalar_ptr(bb), sqrt_den=True)
                dv.lincomb(badd, inv, badd)
                dv.lincomb(xadd, inv, xadd)
                self.x_stored.append(xadd)
                self.b_stored.append(badd)
        return sol

    def solve(self, rhs, x0=None, trans="N"):
        trans = trans.upper()
        _check_trans(trans)
        bd = dv.to_device(rhs)
        x0d = None if x0 is None else dv.to_device(x0)
        if bd.ndim == 2:
            cols = [self._solve1(bd[:, i].contiguous(), None if x0d is None else x0d[:, i].contiguous())
                    for i in range(bd.shape[1])]
            return dv.like_input(torch.stack(cols, dim=1), rhs)
        return dv.like_input(self._solve1(bd.reshape(-1), None if x0d is None else x0d.reshape(-1)), rhs)
