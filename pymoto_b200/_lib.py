"""ctypes binding of libpmb.so (the C ABI declared in include/pmb.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.  The library is
built in-tree by ``pymoto_b200/_build.py`` (``__graft_entry__.build()``) with nvcc for sm_100a.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PMB_LIB_PATH", os.path.join(_HERE, "libpmb.so"))  # PMB_LIB_PATH: diagnostic builds only


class PmbError(RuntimeError):
    pass


class Grid(C.Structure):
    """Mirror of ``pmb_grid`` (include/pmb.h)."""

    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("ndof", C.c_int), ("kz0", C.c_int), ("nzl", C.c_int)]

    def __repr__(self):
        return f"Grid({self.nx}x{self.ny}x{self.nz}, ndof={self.ndof}, planes=[{self.kz0},{self.kz0 + self.nzl}))"


class Coef(C.Structure):
    """Mirror of ``pmb_coef``: c * (*num) / (*den or sqrt(*den))."""

    _fields_ = [("c", C.c_double), ("num", C.c_void_p), ("den", C.c_void_p), ("sqrt_den", C.c_int)]


def coef(c=1.0, num=None, den=None, sqrt_den=False):
    return Coef(float(c), num, den, 1 if sqrt_den else 0)


class Bound(C.Structure):
    """Mirror of ``pmb_bound``: one value for every variable, or a per-variable device array."""

    _fields_ = [("s", C.c_double), ("v", C.c_void_p)]


class MmaVecs(C.Structure):
    """Mirror of ``pmb_mma_vecs`` (device pointers of the n-sized MMA subproblem state)."""

    NAMES = ("x", "xsi", "eta", "xo", "xsio", "etao", "dx", "dxsi", "deta", "low", "upp", "alfa", "beta", "P", "Q")
    _fields_ = [(nm, C.c_void_p) for nm in NAMES]


class ElemOp(C.Structure):
    """Mirror of ``pmb_elem_op``: the matrix-free finest-level operator (host Ke, device s / mask / brick flags, kernel layout)."""

    _fields_ = [("Ke_host", C.c_void_p), ("s", C.c_void_p), ("bcmask", C.c_void_p), ("bcdiagval", C.c_double),
                ("brickflags", C.c_void_p), ("variant", C.c_int)]


MMA_MAXM = 6  # PMB_MMA_MAXM
MAX_LEVELS = 12  # PMB_MAX_LEVELS


class MgLevel(C.Structure):
    """Mirror of ``pmb_mg_level``."""

    _fields_ = [("grid", Grid), ("A", C.c_void_p), ("diag", C.c_void_p), ("u", C.c_void_p), ("u2", C.c_void_p), ("t", C.c_void_p),
                ("rc", C.c_void_p), ("smooth_steps", C.c_int), ("w", C.c_double)]


class MgDesc(C.Structure):
    """Mirror of ``pmb_mg_desc`` (the multigrid hierarchy handed to pmb_vcycle / pmb_pcg_solve)."""

    _fields_ = [("nlevels", C.c_int), ("level", MgLevel * MAX_LEVELS), ("coarse_grid", Grid), ("coarse_inv", C.c_void_p),
                ("coarse_out", C.c_void_p), ("gen", ElemOp)]

PEER_MAX = 16  # PMB_PEER_MAX
PEER_COUNT_MAX = 16  # PMB_PEER_COUNT_MAX


class PeerHalo(C.Structure):
    """Mirror of ``pmb_peer_halo``: mailboxes of the one-launch halo exchange over peer memory."""

    _fields_ = [("box", C.c_void_p), ("box_lo", C.c_void_p), ("box_hi", C.c_void_p), ("ctl", C.c_void_p), ("cap", C.c_longlong)]


class PeerReduce(C.Structure):
    """Mirror of ``pmb_peer_reduce``: tables of the one-launch small all-reduce over peer memory."""

    _fields_ = [("world", C.c_int), ("rank", C.c_int), ("slots", C.c_void_p * PEER_MAX), ("ctl", C.c_void_p)]


_P = C.c_void_p
_LL = C.c_longlong
_D = C.c_double
_I = C.c_int
_G = C.POINTER(Grid)

# name -> (restype, argtypes); every symbol include/pmb.h declares
SIGNATURES = {
    "pmb_last_error": (C.c_char_p, []),
    "pmb_version": (_I, []),
    "pmb_nnz": (_LL, [_G]),
    "pmb_nrows": (_LL, [_G]),
    "pmb_csr_pattern": (_I, [_G, _P, _P, _I, _P]),
    "pmb_assemble": (_I, [_G, _P, _P, _P, _D, _P, _P, _P, _P]),
    "pmb_assemble_sens": (_I, [_G, _P, _P, _P, _P, _P, _I, _P]),
    "pmb_rowstats": (_I, [_G, _P, _P, _P, _P]),
    "pmb_spmv": (_I, [_G, _I, _P, _P, _P, _P, _D, _P, _P, _P, _P, _P]),
    "pmb_spmv_ws_doubles": (_LL, [_G]),
    "pmb_elem_spmv": (_I, [_G, _I, C.POINTER(ElemOp), _P, _P, _P, _D, _P, _P, _P, _P, _P]),
    "pmb_elem_ws_doubles": (_LL, [_G]),
    "pmb_elem_num_variants": (_I, []),
    "pmb_elem_brickflags_bytes": (_LL, [_G, _I]),
    "pmb_elem_brickflags": (_I, [_G, _I, _P, _P, _P]),
    "pmb_elem_autotune_flag_bytes": (_LL, [_G]),
    "pmb_elem_autotune": (_I, [_G, C.POINTER(ElemOp), _P, _P, _P, _P, _P, _I, _P, C.POINTER(C.c_int), _P]),
    "pmb_sym_doubles": (_LL, [_G]),
    "pmb_sym_pack": (_I, [_G, _P, _P, _P, _P]),
    "pmb_sym_spmv": (_I, [_G, _I, _P, _P, _P, _P, _D, _P, _P, _P, _P, _P]),
    "pmb_ws_doubles": (_LL, []),
    "pmb_smooth0": (_I, [_LL, _D, _P, _P, _P, _P]),
    "pmb_restrict": (_I, [_G, _G, _P, _P, _P]),
    "pmb_prolong_add": (_I, [_G, _G, _P, _P, _P]),
    "pmb_galerkin": (_I, [_G, _G, _P, _P, _P, _P]),
    "pmb_galerkin_ws_doubles": (_LL, [_G]),
    "pmb_galerkin_cols": (_I, [_G, _G, _P, _P, _P]),
    "pmb_galerkin_rows": (_I, [_G, _G, _P, _P, _P]),
    "pmb_galerkin_direct": (_I, [_G, _G, _P, _P, _P, _P, _P, _P]),
    "pmb_scatter_add": (_I, [_LL, _P, _P, _P, _P]),
    "pmb_densify": (_I, [_G, _P, _P, _P]),
    "pmb_dense_invert": (_I, [_I, _P, _P, _P, _P]),
    "pmb_dense_invert_ws_doubles": (_LL, [_I]),
    "pmb_dense_gemv": (_I, [_I, _P, _P, _P, _P]),
    "pmb_dots": (_I, [_LL, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "pmb_lincomb": (_I, [_LL, _P, Coef, _P, Coef, _P, _P]),
    "pmb_cg_xr_update": (_I, [_LL, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "pmb_bc_split": (_I, [_LL, _P, _P, _P, _P, _P, _P]),
    "pmb_diag_mask": (_I, [_LL, _P, _P, _P, _P]),
    "pmb_mask_zero": (_I, [_LL, _P, _P, _P, _P]),
    "pmb_filter_apply": (_I, [_G, _I, _I, _I, _P, _P, _P, _P, _P]),
    "pmb_pad_gather": (_I, [_I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "pmb_pad_scatter": (_I, [_I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "pmb_stencil_corr": (_I, [_I, _I, _I, _P, _I, _I, _I, _P, _I, _I, _I, _P, _I, _I, _I, _P]),
    "pmb_vec_div": (_I, [_LL, _P, _P, _P, _P]),
    "pmb_halo_copy2": (_I, [_LL, _P, _P, _P, _P, _P]),
    "pmb_peer_halo_box_doubles": (_LL, [_LL]),
    "pmb_peer_reduce_table_doubles": (_LL, [_I]),
    "pmb_peer_halo_exchange": (_I, [C.POINTER(PeerHalo), _LL, _P, _P, _P, _P, _P]),
    "pmb_peer_allreduce": (_I, [C.POINTER(PeerReduce), _P, _I, _I, _P]),
    "pmb_comm_unique_id": (_I, [_P]),
    "pmb_comm_init": (_I, [_P, _I, _I, C.POINTER(C.c_void_p)]),
    "pmb_comm_destroy": (_I, [_P]),
    "pmb_comm_rank": (_I, [_P]),
    "pmb_comm_size": (_I, [_P]),
    "pmb_halo_exchange": (_I, [_P, _P, _LL, _LL, _LL, _I, _I, _P]),
    "pmb_allreduce": (_I, [_P, _P, _LL, _P]),
    "pmb_oc_candidate": (_I, [_LL, _P, _P, _D, _D, _D, _D, _P, _P, _P, _P]),
    "pmb_mma_ws_doubles": (_LL, []),
    "pmb_mma_asymptotes": (_I, [_LL, _P, _P, _P, _D, _D, _D, _P, _P]),
    "pmb_mma_setup": (_I, [_LL, _I, _P, C.POINTER(C.c_void_p), _P, Bound, Bound, Bound, _D, C.POINTER(C.c_double), _I,
                           C.POINTER(MmaVecs), _P, _P, _P]),
    "pmb_mma_residual": (_I, [_LL, _I, C.POINTER(MmaVecs), C.POINTER(C.c_double), _D, _P, _P, _P]),
    "pmb_mma_newton_sums": (_I, [_LL, _I, C.POINTER(MmaVecs), C.POINTER(C.c_double), _D, _P, _P, _P]),
    "pmb_mma_newton_dir": (_I, [_LL, _I, C.POINTER(MmaVecs), C.POINTER(C.c_double), C.POINTER(C.c_double), _D, _P, _P, _P]),
    "pmb_mma_linesearch": (_I, [_LL, _I, C.POINTER(MmaVecs), C.POINTER(C.c_double), _D, _D, _P, _P, _P]),
    "pmb_mma_gcmma_rho": (_I, [_LL, _I, C.POINTER(C.c_void_p), Bound, Bound, _P, _P, _P]),
    "pmb_mma_gcmma_estimate": (_I, [_LL, _I, C.POINTER(MmaVecs), _P, Bound, Bound, _P, _P, _P]),
    "pmb_probe_fp64_out_doubles": (_LL, []),
    "pmb_probe_fp64": (_I, [_I, _I, _P, C.POINTER(C.c_double), _P]),
    "pmb_pack_f32": (_I, [_LL, _I, _I, _P, _P, _P]),
    "pmb_vcycle": (_I, [C.POINTER(MgDesc), _P, _P, _P]),
    "pmb_pcg_solve": (_I, [C.POINTER(MgDesc), _P, _P, _P, _P, _P, _D, _I, _I, _P, _P, _P, C.POINTER(C.c_int),
                           C.POINTER(C.c_double), _P]),
    "pmb_pcg_plan_create": (_I, [C.POINTER(MgDesc), _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(C.c_void_p)]),
    "pmb_pcg_plan_solve": (_I, [_P, _D, _I, _I, C.POINTER(C.c_int), C.POINTER(C.c_double), _P]),
    "pmb_pcg_plan_graph_replays": (_LL, [_P]),
    "pmb_pcg_plan_destroy": (_I, [_P]),
    "pmb_simp": (_I, [_LL, _D, _I, _P, _P, _P]),
    "pmb_simp_bwd": (_I, [_LL, _D, _I, _P, _P, _P, _P]),
}

SPMV, RESIDUAL, JACOBI = 0, 1, 2

_lib = None
launch_count = 0  # CUDA kernels launched by libpmb through this binding (bench.py reports it as gpu_launches)
call_stats = {}   # (entry point, detail) -> number of calls; detail = (nx, mode) for pmb_spmv


def _kernels_launched(name, args):
    """How many kernels one C-ABI call launches (see the .cu sources)."""
    if name == "pmb_spmv":
        return 2 if args[9] is not None else 1  # + reduce_triples_kernel when the fused dots are requested
    if name == "pmb_elem_spmv" or name == "pmb_sym_spmv":
        return 2 if args[9] is not None else 1
    if name == "pmb_elem_autotune":
        return 8 * load().pmb_elem_num_variants()  # every variant: 2 warm-up + 6 timed launches
    if name == "pmb_probe_fp64":
        return 4
    if name.startswith("pmb_comm_") or name in ("pmb_halo_exchange", "pmb_allreduce"):
        return 0  # NCCL's own kernels
    if name == "pmb_galerkin":
        return 2  # column-collapse + row-collapse passes
    if name == "pmb_dense_invert":
        return 1  # one cooperative kernel (the three-kernel fallback: 3 per 32 pivots)
    return 1


def load():
    """Load libpmb.so; raises PmbError when it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PmbError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a). pymoto_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


profile_times = None  # set to {} to time every call with CUDA events (synchronising; diagnostics only)


def _stat_key(name, args):
    """(entry point, detail): detail = (nx, mode) for the operator kernels, (nx, None) for other grid-first calls, (n, None)
    for vector calls -- enough for bench.py to attribute launches and algorithmic bytes to multigrid levels."""
    a0 = args[0] if args else None
    if name in ("pmb_spmv", "pmb_elem_spmv", "pmb_sym_spmv"):
        return (name, (a0.nx, args[1]))
    if isinstance(a0, Grid):
        return (name, (a0.nx, None))
    if isinstance(a0, int):
        return (name, (a0, None))
    return (name, None)


def call(name, *args):
    """Call an int-returning entry point and raise PmbError with pmb_last_error() on failure."""
    global launch_count
    lib = load()
    key = _stat_key(name, args)
    if profile_times is not None:
        import torch

        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        e1.synchronize()
        t = profile_times.setdefault(key, [0, 0.0])
        t[0] += 1
        t[1] += e0.elapsed_time(e1)
    else:
        rc = getattr(lib, name)(*args)
    launch_count += _kernels_launched(name, args)
    call_stats[key] = call_stats.get(key, 0) + 1
    if rc != 0:
        raise PmbError(f"{name} failed: {lib.pmb_last_error().decode()}")


def query(name, *args):
    lib = load()
    v = getattr(lib, name)(*args)
    if v < 0:
        raise PmbError(f"{name} failed: {lib.pmb_last_error().decode()}")
    return v
