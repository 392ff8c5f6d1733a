"""CPU-only tests: the C-ABI library loads and exports every symbol include/pmb.h declares, the host-side mirror of
the reference interface behaves (no compute calls without a GPU), and the product path fails loudly without CUDA."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge

    ge.build()
    from pymoto_b200 import _lib

    return _lib


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "pmb.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(pmb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    cdll = ctypes.CDLL(lib.LIB_PATH)
    for name in declared:
        assert hasattr(cdll, name), f"{name} declared in pmb.h but not exported"
    assert declared == set(lib.SIGNATURES), declared ^ set(lib.SIGNATURES)
    assert lib.load().pmb_version() >= 100


def test_size_queries_and_argument_validation(lib):
    """Host-side entry points that launch nothing: closed-form nnz and error reporting through pmb_last_error."""
    from oracle import Grid
    import oracle

    for shape, ndof in [((6, 4, 4), 3), ((12, 8, 0), 2), ((8, 8, 8), 1), ((64, 32, 32), 3), ((256, 128, 128), 3)]:
        g = lib.Grid(shape[0], shape[1], shape[2], ndof, 0, shape[2] + 1)
        nnz = lib.query("pmb_nnz", g)
        NX, NY, NZ = shape[0] + 1, shape[1] + 1, shape[2] + 1
        assert nnz == ndof * ndof * (3 * NX - 2) * (3 * NY - 2) * (3 * NZ - 2)
        assert lib.query("pmb_nrows", g) == NX * NY * NZ * ndof
        if np.prod(shape[:2]) * max(shape[2], 1) < 2000:
            ip, ix = oracle.assembly.pattern_closed_form(Grid(*shape), ndof)
            assert nnz == ix.size == ip[-1]
    # slabs partition the matrix
    full = lib.query("pmb_nnz", lib.Grid(16, 8, 8, 3, 0, 9))
    parts = [lib.query("pmb_nnz", lib.Grid(16, 8, 8, 3, k0, n)) for k0, n in [(0, 4), (4, 2), (6, 3)]]
    assert sum(parts) == full
    assert lib.load().pmb_nnz(lib.Grid(0, 4, 4, 3, 0, 5)) == -1
    assert b"invalid grid" in lib.load().pmb_last_error()
    with pytest.raises(lib.PmbError):
        lib.query("pmb_nnz", lib.Grid(4, 4, 4, 5, 0, 5))
    with pytest.raises(lib.PmbError):
        lib.call("pmb_csr_pattern", lib.Grid(4, 4, 4, 3, 0, 5), None, None, 32, None)  # NULL outputs rejected before launch


def test_product_path_fails_loudly_without_cuda():
    import torch
    import pymoto_b200 as pmb

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    dom = pmb.VoxelDomain(4, 4, 4)
    for ctor in (lambda: pmb.DensityFilter(dom), lambda: pmb.AssembleStiffness(dom), lambda: pmb.AssemblePoisson(dom)):
        with pytest.raises(pmb.PmbError):
            ctor()
    with pytest.raises(TypeError):
        pmb.LinSolve()(np.eye(3), np.ones(3))  # only DeviceCSR is accepted: no scipy / CPU route


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pymoto_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_domain_matches_oracle_grid():
    import pymoto_b200 as pmb
    from oracle import Grid

    for shape in [(5, 4, 3), (7, 3, 0)]:
        d, g = pmb.VoxelDomain(*shape), Grid(*shape)
        assert (d.nel, d.nnodes, d.dim, d.elemnodes) == (g.nel, g.nnodes, g.dim, g.elemnodes)
        assert np.array_equal(d.conn, g.conn())
        assert np.array_equal(d.get_dofconnectivity(2), g.dofconn(2))
        assert np.array_equal(d.nodes, g.nodes3d())
        n = np.arange(d.nnodes)
        ijk = d.get_node_indices(n)
        k = ijk[2] if d.dim == 3 else 0
        assert np.array_equal(d.get_nodenumber(ijk[0], ijk[1], k), n)  # reference tests/test_domain.py round trip
        pos = np.array([0.1, -0.2, 0.3])
        assert abs(d.eval_shape_fun(pos).sum() - 1.0) < 1e-14  # partition of unity
        assert np.array_equal(d.eval_shape_fun_der(pos), g.shape_fun_der(pos))
    assert pmb.DomainDefinition is pmb.VoxelDomain
    assert d.get_dofnumber(np.array([1, 2]), ndof=3).tolist() == [[3, 4, 5], [6, 7, 8]]


def test_element_matrices_match_oracle():
    """Host-side Gauss integration of the product equals the oracle's (itself bit-equal to the reference's)."""
    import pymoto_b200.assembly as pa
    import pymoto_b200 as pmb
    import oracle
    from oracle import Grid

    for shape in [(3, 2, 2), (3, 2, 0)]:
        d, g = pmb.VoxelDomain(*shape, unitx=1.0, unity=0.5, unitz=2.0), Grid(*shape, unitx=1.0, unity=0.5, unitz=2.0)
        D = pa._elasticity_matrix(1.0, 0.3, "3d" if d.dim == 3 else "strain")
        Ke = np.zeros((d.elemnodes * d.dim,) * 2)
        w = np.prod(d.element_size[: d.dim] / 2) * (d.element_size[2] if d.dim == 2 else 1.0)
        for pos in pa._gauss_points(d):
            B = pa._strain_displacement(d.eval_shape_fun_der(pos))
            Ke += w * B.T @ D @ B
        assert np.array_equal(Ke, oracle.assembly.stiffness_element(g))


def test_module_runtime_protocol():
    """Signal/Module/Network stand-ins: connection, response order, reverse-order back-propagation, reset."""
    import pymoto_b200 as pmb

    if pmb.HAVE_PYMOTO:
        pytest.skip("real pymoto runtime in use")

    class Scale(pmb.Module):
        def __init__(self, a):
            self.a = a

        def __call__(self, x):
            return self.a * x

        def _sensitivity(self, dy):
            return self.a * dy

    class Sum(pmb.Module):
        def __call__(self, x, y):
            return x.sum() + y.sum()

        def _sensitivity(self, dc):
            x, y = self.get_input_states()
            return dc * np.ones_like(x), dc * np.ones_like(y)

    assert Scale(2.0)(np.ones(3)).tolist() == [2.0, 2.0, 2.0]  # plain-function use
    sx = pmb.Signal("x", state=np.arange(3.0))
    with pmb.Network() as fn:
        sy = Scale(3.0)(sx)
        sc = Sum()(sy, sx)
    assert len(fn.mods) == 2 and sc.state == 12.0
    sx.state = 2 * np.ones(3)
    fn.response()
    assert sc.state == 24.0
    sc.sensitivity = 1.0
    fn.sensitivity()
    assert sx.sensitivity.tolist() == [4.0, 4.0, 4.0]
    fn.reset()
    assert sx.sensitivity is None and sy.sensitivity is None


def test_header_and_binding_agree_on_arity():
    """Every prototype in include/pmb.h has the same number of parameters as its ctypes signature in _lib.SIGNATURES."""
    from pymoto_b200 import _lib

    header = open(os.path.join(ROOT, "include", "pmb.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    protos = re.findall(r"\b(pmb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S)
    assert len(protos) >= 30
    for name, params in protos:
        params = params.strip()
        n = 0 if params in ("", "void") else len([p for p in params.split(",") if p.strip()])
        assert name in _lib.SIGNATURES, name
        assert n == len(_lib.SIGNATURES[name][1]), (name, n, len(_lib.SIGNATURES[name][1]))


def test_filterconv_axis_maps_match_reference_padding():
    """The per-axis index maps of FilterConv reproduce the reference's np.pad sequence (filter.py:99-160), including
    mixed boundary types; checked against the live reference where it is available."""
    from pymoto_b200.filter import FilterConv
    from _refimport import import_reference

    m, c = FilterConv._axis_map(6, 2, "symmetric", "symmetric")
    assert m.tolist() == [1, 0, 0, 1, 2, 3, 4, 5, 5, 4]
    m, c = FilterConv._axis_map(6, 2, 0.5, "wrap")
    assert m.tolist() == [-1, -1, 0, 1, 2, 3, 4, 5, 0, 1] and c[:2].tolist() == [0.5, 0.5]
    assert FilterConv._axis_map(5, 0, "edge", "edge")[0].tolist() == [0, 1, 2, 3, 4]
    with pytest.raises(ValueError):
        FilterConv._axis_map(5, 1, "mirror", "edge")
    pym = import_reference()
    if pym is None:
        return
    d = pym.VoxelDomain(7, 1, 1)
    for bc0, bc1 in [("symmetric", "edge"), ("edge", "wrap"), ("wrap", "symmetric"), ("wrap", "wrap"), ("edge", "edge")]:
        ref = pym.FilterConv(d, weights=np.ones((5, 1, 1)), xmin_bc=bc0, xmax_bc=bc1)
        assert ref.el3d_pad[:, 0, 0].tolist() == FilterConv._axis_map(7, 2, bc0, bc1)[0].tolist(), (bc0, bc1)


def test_ctypes_struct_mirrors_match_the_header_layout(tmp_path):
    """sizeof / offsetof of every struct in include/pmb.h as the C compiler lays it out vs the ctypes mirrors in
    pymoto_b200/_lib.py (a silent mismatch would corrupt arguments passed by value or by pointer)."""
    import subprocess

    from pymoto_b200 import _lib

    src = tmp_path / "layout.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "pmb.h"
#define S(t) printf(#t " %zu\n", sizeof(t))
#define O(t, f) printf(#t "." #f " %zu\n", offsetof(t, f))
int main(void) {
  S(pmb_grid); S(pmb_coef); S(pmb_bound); S(pmb_mma_vecs); S(pmb_mg_level); S(pmb_mg_desc); S(pmb_elem_op); S(pmb_peer_halo); S(pmb_peer_reduce);
  O(pmb_coef, sqrt_den); O(pmb_bound, v); O(pmb_mma_vecs, Q); O(pmb_mg_level, A); O(pmb_mg_level, smooth_steps);
  O(pmb_mg_level, w); O(pmb_mg_desc, level); O(pmb_mg_desc, coarse_grid); O(pmb_mg_desc, coarse_inv); O(pmb_mg_desc, gen);
  O(pmb_elem_op, bcdiagval); O(pmb_elem_op, brickflags); O(pmb_elem_op, variant);
  O(pmb_peer_halo, box_hi); O(pmb_peer_halo, ctl); O(pmb_peer_halo, cap); O(pmb_peer_reduce, slots); O(pmb_peer_reduce, ctl);
  printf("PMB_MMA_MAXM %d\nPMB_MAX_LEVELS %d\nPMB_PEER_MAX %d\nPMB_PEER_COUNT_MAX %d\n", PMB_MMA_MAXM, PMB_MAX_LEVELS, PMB_PEER_MAX, PMB_PEER_COUNT_MAX);
  return 0;
}
''')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(line.rsplit(" ", 1) for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    mirrors = {"pmb_grid": _lib.Grid, "pmb_coef": _lib.Coef, "pmb_bound": _lib.Bound, "pmb_mma_vecs": _lib.MmaVecs,
               "pmb_mg_level": _lib.MgLevel, "pmb_mg_desc": _lib.MgDesc, "pmb_elem_op": _lib.ElemOp,
               "pmb_peer_halo": _lib.PeerHalo, "pmb_peer_reduce": _lib.PeerReduce}
    for name, cls in mirrors.items():
        assert int(got[name]) == ctypes.sizeof(cls), (name, got[name], ctypes.sizeof(cls))
    for key, val in got.items():
        if "." in key:
            st, field = key.split(".")
            assert int(val) == getattr(mirrors[st], field).offset, (key, val)
    assert int(got["PMB_MMA_MAXM"]) == _lib.MMA_MAXM and int(got["PMB_MAX_LEVELS"]) == _lib.MAX_LEVELS
    assert int(got["PMB_PEER_MAX"]) == _lib.PEER_MAX and int(got["PMB_PEER_COUNT_MAX"]) == _lib.PEER_COUNT_MAX


def test_direct_coarse_operator_tables_vs_oracle():
    """Host set-up of the direct level-1 Galerkin build (pymoto_b200/coarse.py): eight child tables, Dirichlet-pattern
    tables, per-coarse-element lookup and the bc diagonal term, evaluated in plain numpy, equal the oracle's R^T A R --
    whole grid and as coarse slabs (multi-GPU row ranges)."""
    import scipy.sparse as sps

    import oracle.assembly as oasm
    import oracle.solvers as osol
    from oracle import Grid
    from pymoto_b200 import coarse

    rng = np.random.default_rng(0)
    for dims, ndof in [((4, 4, 4), 3), ((6, 4, 4), 1)]:
        g, gc = Grid(*dims), Grid(*(d // 2 for d in dims))
        Ke = rng.standard_normal((8 * ndof,) * 2)
        Ke = Ke + Ke.T + 8 * np.eye(8 * ndof)
        bc = np.unique(rng.integers(0, g.nnodes * ndof, 9))
        asm = oasm.Assembler(g, Ke, bc=bc)
        s = rng.random(g.nel)
        R = osol.prolongation_matrix(g, gc, ndof)
        ref = (R.T @ asm(s) @ R).tocsr()
        ref.sort_indices()
        tabs = coarse.build_tables(Ke, ndof, dims, bc, asm.bcdiagval)
        assert tabs["Gtab"].shape[0] > 8 and tabs["cidx"].max() >= 0
        data = coarse.emulate(tabs, ndof, dims, s)
        ip, ix = oasm.pattern_closed_form(gc, ndof)
        mine = sps.csr_matrix((data, ix, ip), shape=ref.shape)
        assert abs(mine - ref).max() <= 1e-13 * abs(ref).max()
        plane = (gc.nelx + 1) * (gc.nely + 1) * ndof
        for k0c, k1c in [(0, 1), (1, dims[2] // 2 + 1)]:
            t2 = coarse.build_tables(Ke, ndof, dims, bc, asm.bcdiagval, k0c, k1c)
            assert np.array_equal(coarse.emulate(t2, ndof, dims, s, k0c, k1c), data[ip[k0c * plane]:ip[k1c * plane]])
