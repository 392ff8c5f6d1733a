"""Pure C-ABI check (run as a script on a GPU box; NO torch in this process):

    python tests/c_abi_check.py            # single GPU: assemble + Galerkin hierarchy + pmb_pcg_solve at 32x16x16 vs the oracle
    python tests/c_abi_check.py comm R N F # rank R of N: pmb_comm_init / pmb_halo_exchange / pmb_allreduce (id exchanged via file F)

Device memory comes from cudaMalloc through ctypes, every call goes through libpmb.so's extern "C" entry points with plain
pointers -- what a non-Python, non-torch host would do.  Prints "[c_abi_check] OK" on success.
"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class Cuda:
    def __init__(self):
        self.rt = None
        for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
            try:
                self.rt = C.CDLL(name)
                break
            except OSError:
                pass
        assert self.rt is not None, "libcudart not found"
        self.rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        self.rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        self.rt.cudaMemset.argtypes = [C.c_void_p, C.c_int, C.c_size_t]
        self.rt.cudaSetDevice.argtypes = [C.c_int]

    def ok(self, rc, what):
        assert rc == 0, f"{what}: cudaError {rc}"

    def malloc(self, nbytes, zero=False):
        p = C.c_void_p()
        self.ok(self.rt.cudaMalloc(C.byref(p), max(int(nbytes), 16)), "cudaMalloc")
        if zero:
            self.ok(self.rt.cudaMemset(p, 0, max(int(nbytes), 16)), "cudaMemset")
        return p

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        p = self.malloc(arr.nbytes + 16)  # + one 16-byte granule (libpmb reads whole granules of matrix / vector tails)
        self.ok(self.rt.cudaMemcpy(p, arr.ctypes.data, arr.nbytes, 1), "cudaMemcpy H2D")
        return p

    def download(self, p, n, dtype=np.float64):
        out = np.empty(n, dtype=dtype)
        self.ok(self.rt.cudaMemcpy(out.ctypes.data, p, out.nbytes, 2), "cudaMemcpy D2H")
        return out

    def sync(self):
        self.ok(self.rt.cudaDeviceSynchronize(), "cudaDeviceSynchronize")


def load_lib():
    """The ctypes binding and the build recipe loaded BY FILE, so that the pymoto_b200 package (which imports torch) is not."""
    import importlib.util

    mods = {}
    for name in ("_build", "_lib"):
        spec = importlib.util.spec_from_file_location("pmb" + name, os.path.join(ROOT, "pymoto_b200", name + ".py"))
        mods[name] = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mods[name])
    mods["_build"].build()  # nvcc, only if libpmb.so is missing or stale
    return mods["_lib"]


def check(lib, rc, name):
    assert rc == 0, f"{name}: {lib.load().pmb_last_error().decode()}"


def solve_single_gpu():
    import oracle.assembly as oasm
    import oracle.solvers as osol
    from oracle import Grid
    from oracle.chain import cantilever

    assert "torch" not in sys.modules
    lib = load_lib()
    L = lib.load()
    cu = Cuda()
    cu.ok(cu.rt.cudaSetDevice(0), "cudaSetDevice")
    dims = [(32, 16, 16), (16, 8, 8), (8, 4, 4)]
    grids = [lib.Grid(nx, ny, nz, 3, 0, nz + 1) for nx, ny, nz in dims]
    gr = Grid(*dims[0])
    ndof, bc, f = cantilever(gr)
    Ke = oasm.stiffness_element(gr)
    s = 0.1 + 0.9 * np.random.default_rng(3).random(gr.nel)
    n = [L.pmb_nrows(C.byref(g)) for g in grids]
    nnz = [L.pmb_nnz(C.byref(g)) for g in grids]
    mask = np.zeros(n[0], np.uint8)
    mask[bc] = 1
    bcdiag = float(Ke.max())
    d_s = cu.upload(np.concatenate([np.zeros(2 * 32 * 16), s]))  # two halo layers in front (unused on one GPU)
    d_s = C.c_void_p(d_s.value + 8 * 2 * 32 * 16)
    d_mask = cu.upload(mask)
    A = [cu.malloc(8 * (z + 2), zero=True) for z in nnz]
    diag = [cu.malloc(8 * m) for m in n[:2]]
    Keh = np.ascontiguousarray(Ke.ravel())
    check(lib, L.pmb_assemble(C.byref(grids[0]), Keh.ctypes.data, d_s, d_mask, bcdiag, A[0], None, None, None), "pmb_assemble")
    work = cu.malloc(8 * (L.pmb_galerkin_ws_doubles(C.byref(grids[0])) + 16))
    for l in range(2):
        check(lib, L.pmb_galerkin(C.byref(grids[l]), C.byref(grids[l + 1]), A[l], A[l + 1], work, None), "pmb_galerkin")
        check(lib, L.pmb_rowstats(C.byref(grids[l]), A[l], diag[l], None, None), "pmb_rowstats")
    dense = cu.malloc(8 * n[2] * n[2])
    check(lib, L.pmb_densify(C.byref(grids[2]), A[2], dense, None), "pmb_densify")
    info = cu.malloc(16, zero=True)
    scratch = cu.malloc(8 * L.pmb_dense_invert_ws_doubles(n[2]))
    check(lib, L.pmb_dense_invert(n[2], dense, scratch, info, None), "pmb_dense_invert")
    cu.sync()
    assert cu.download(info, 1, np.int32)[0] == 0, "coarsest operator not positive definite"

    desc = lib.MgDesc()
    desc.nlevels = 2
    pad = 8 * (33 * 17 * 3)  # vectors the operators are applied to carry one node plane of padding on both sides
    for l in range(2):
        lv = desc.level[l]
        lv.grid, lv.A, lv.diag = grids[l], A[l], diag[l]
        for nm in ("u", "u2", "t"):
            setattr(lv, nm, cu.malloc(8 * n[l] + 2 * pad, zero=True).value + pad)
        lv.rc = cu.malloc(8 * n[l + 1], zero=True).value
        lv.smooth_steps, lv.w = 5, 0.5
    desc.coarse_grid, desc.coarse_inv, desc.coarse_out = grids[2], dense, cu.malloc(8 * n[2])
    for matrix_free in (False, True):
        if matrix_free:
            desc.gen = lib.ElemOp(Keh.ctypes.data, d_s, d_mask, bcdiag, None, 0)
        b = f.copy()
        b[bc] = 0.0
        d_b = cu.upload(b)
        vec = lambda: C.c_void_p(cu.malloc(8 * n[0] + 2 * pad, zero=True).value + pad)  # noqa: E731
        x, r, q, p = vec(), vec(), vec(), vec()
        scal = cu.malloc(8 * 16, zero=True)
        ws_red = cu.malloc(8 * L.pmb_ws_doubles(), zero=True)
        ws_spmv = cu.malloc(8 * max(L.pmb_spmv_ws_doubles(C.byref(grids[0])), L.pmb_elem_ws_doubles(C.byref(grids[0]))))
        iters, relres = C.c_int(0), C.c_double(0.0)
        t0 = time.perf_counter()
        check(lib, L.pmb_pcg_solve(C.byref(desc), d_b, x, r, q, p, 1e-8, 200, 50, scal, ws_red, ws_spmv, C.byref(iters), C.byref(relres),
                                   None), "pmb_pcg_solve")
        cu.sync()
        dt = time.perf_counter() - t0
        u = cu.download(x, n[0])
        # oracle: the reference's assembly + CG(GMG) restated in numpy / scipy on the same inputs
        K = oasm.Assembler(gr, Ke, bc=bc)(s)
        res = np.linalg.norm(K @ u - b) / np.linalg.norm(b)
        mgs = osol.make_gmg_chain(gr, min_size=8)
        cg = osol.CG(mgs[0], tol=1e-8)
        cg.update(K)
        u_ref = cg.solve(b)
        err = np.linalg.norm(u - u_ref) / np.linalg.norm(u_ref)
        print(f"[c_abi_check] pmb_pcg_solve ({'matrix-free level 0' if matrix_free else 'CSR on every level'}): {iters.value} iterations "
              f"(oracle {cg.iterations}), relres {relres.value:.2e}, |K u - b| / |b| = {res:.2e} on the oracle matrix, "
              f"|u - u_oracle| / |u_oracle| = {err:.2e}, {1e3 * dt:.1f} ms")
        assert relres.value <= 1e-8 and res <= 2e-8 and err <= 1e-6 and abs(iters.value - cg.iterations) <= 1
    assert "torch" not in sys.modules
    print("[c_abi_check] OK")


def comm_check(rank, nranks, idfile):
    lib = load_lib()
    L = lib.load()
    cu = Cuda()
    cu.ok(cu.rt.cudaSetDevice(rank), "cudaSetDevice")
    idbuf = (C.c_char * 128)()
    if rank == 0:
        check(lib, L.pmb_comm_unique_id(idbuf), "pmb_comm_unique_id")
        with open(idfile + ".tmp", "wb") as fh:
            fh.write(bytes(idbuf))
        os.replace(idfile + ".tmp", idfile)
    else:
        for _ in range(600):
            if os.path.exists(idfile):
                break
            time.sleep(0.1)
        idbuf = (C.c_char * 128).from_buffer_copy(open(idfile, "rb").read())
    comm = C.c_void_p()
    check(lib, L.pmb_comm_init(idbuf, rank, nranks, C.byref(comm)), "pmb_comm_init")
    assert L.pmb_comm_rank(comm) == rank and L.pmb_comm_size(comm) == nranks
    plane, own = 1000, 5000
    host = np.full(own + 2 * plane, -1.0)
    host[plane:plane + own] = rank * 10000.0 + np.arange(own)
    d = cu.upload(host)
    check(lib, L.pmb_halo_exchange(comm, d, plane, own, plane, 1, 1, None), "pmb_halo_exchange")
    red = cu.upload(np.array([rank + 1.0, 2.0 * rank, 1.0, 0.5]))
    check(lib, L.pmb_allreduce(comm, red, 4, None), "pmb_allreduce")
    cu.sync()
    got = cu.download(d, host.size)
    if rank > 0:
        assert np.array_equal(got[:plane], (rank - 1) * 10000.0 + np.arange(own - plane, own)), "lower halo"
    else:
        assert np.all(got[:plane] == -1.0)
    if rank < nranks - 1:
        assert np.array_equal(got[plane + own:], (rank + 1) * 10000.0 + np.arange(plane)), "upper halo"
    else:
        assert np.all(got[plane + own:] == -1.0)
    r = cu.download(red, 4)
    assert np.allclose(r, [nranks * (nranks + 1) / 2, nranks * (nranks - 1), nranks, 0.5 * nranks])
    check(lib, L.pmb_comm_destroy(comm), "pmb_comm_destroy")
    print(f"[c_abi_check] comm rank {rank}/{nranks} OK")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "comm":
        comm_check(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
    else:
        solve_single_gpu()
