"""GPU parity tests: every kernel of the hot path against the CPU oracle and the golden fixtures generated from the
unmodified reference.  Bit-exact where the contract says so (CSR pattern, assembled values, filter, transfer
operators); tolerances written next to each floating-point comparison (north star: values rtol 1e-12, relative
residual <= 1e-8, compliance within 1e-6 relative)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle
from oracle import Grid
from oracle.chain import ComplianceProblem, cantilever
from _golden import CASES, DESIGNS, load, digest, force_vector


@pytest.fixture(scope="module")
def pmb():
    import torch
    import pymoto_b200 as pmb

    assert torch.cuda.is_available()
    return pmb


def _asm_for(pmb, g, kind, bc):
    nx, ny, nz = (int(v) for v in g["shape"])
    dom = pmb.VoxelDomain(nx, ny, nz)
    if kind == "heatsink":
        return dom, pmb.AssemblePoisson(dom, bc=bc)
    return dom, pmb.AssembleStiffness(dom, bc=bc)


# ------------------------------------------------------------------------------------------------ assembly
@pytest.mark.parametrize("case", CASES)
def test_assembly_bit_exact_vs_golden(pmb, case):
    g = load(case)
    dom, asm = _asm_for(pmb, g, str(g["kind"]), g["bc"])
    assert np.array_equal(asm.elmat[0], g["Ke"])  # host element matrix identical to the reference's
    assert float(asm.bcdiagval) == float(g["bcdiagval"])
    xmin = float(g["xmin"])
    for dname in DESIGNS:
        s = xmin + (1.0 - xmin) * g[dname + "_y"] ** 3
        K = asm(s)
        assert K.nnz == int(g["nnz"])
        assert digest(K.indptr.cpu().numpy()) == str(g["indptr_sha256"])
        assert digest(K.indices.cpu().numpy()) == str(g["indices_sha256"])
        assert K.indptr.cpu().numpy().dtype == np.int32
        assert digest(K.data.cpu().numpy()) == str(g[dname + "_data_sha256"])
        assert np.array_equal(K.diagonal(), g[dname + "_diag"])
        if dname == "random" and "indptr" in g.files:
            csr = K.tocsr()
            assert np.array_equal(csr.indptr, g["indptr"]) and np.array_equal(csr.indices, g["indices"])
            assert np.array_equal(csr.data, g["data_random"])


@pytest.mark.parametrize("shape,ndof", [((5, 3, 2), 3), ((7, 4, 0), 2), ((3, 5, 4), 1), ((1, 1, 1), 3), ((2, 1, 0), 2),
                                        ((33, 9, 5), 3), ((40, 3, 3), 1), ((2, 2, 0), 1)])
def test_assembly_vs_oracle_ragged(pmb, shape, ndof):
    rng = np.random.default_rng(42)
    gr = Grid(*shape)
    dom = pmb.VoxelDomain(*shape)
    Ke = rng.standard_normal((gr.elemnodes * ndof, gr.elemnodes * ndof))  # general (unsymmetric) element matrix
    bc = np.unique(rng.integers(0, gr.nnodes * ndof, 7))
    x = rng.random(gr.nel)
    Ko = oracle.assembly.Assembler(gr, Ke, bc=bc, bcdiagval=3.25, closed_form=False)(x)
    K = pmb.AssembleGeneral(dom, Ke, bc=bc, bcdiagval=3.25)(x).tocsr()
    assert np.array_equal(K.indptr, Ko.indptr) and np.array_equal(K.indices, Ko.indices)
    assert np.array_equal(K.data, Ko.data)
    # no bc
    Ko = oracle.assembly.Assembler(gr, Ke, closed_form=False)(x)
    K = pmb.AssembleGeneral(dom, Ke)(x).tocsr()
    assert np.array_equal(K.data, Ko.data)


def test_assembly_single_element_and_errors(pmb):
    """reference tests/test_assembly.py:22-35 (one element == Ke) and the constructor's error behaviour."""
    dom = pmb.VoxelDomain(1, 1, 1)
    asm = pmb.AssembleStiffness(dom)
    assert np.array_equal(asm(np.array([1.0])).toarray(), asm.elmat[0])
    with pytest.raises(ValueError):
        asm(np.ones(3))
    with pytest.raises(ValueError):
        pmb.AssembleGeneral(dom, np.ones((7, 7)))


# ------------------------------------------------------------------------------------------------ operator kernels
@pytest.mark.parametrize("shape,ndof", [((6, 4, 4), 3), ((12, 8, 0), 2), ((8, 8, 8), 1), ((33, 9, 5), 3), ((17, 6, 0), 1),
                                        ((130, 5, 3), 1), ((20, 18, 3), 2)])
def test_spmv_residual_jacobi_rowstats(pmb, shape, ndof):
    import torch
    from pymoto_b200 import _lib, device as dv

    rng = np.random.default_rng(1)
    gr = Grid(*shape)
    dom = pmb.VoxelDomain(*shape)
    Ke = rng.standard_normal((gr.elemnodes * ndof,) * 2)
    Ke = Ke + Ke.T + 8 * np.eye(Ke.shape[0])
    bc = np.unique(rng.integers(0, gr.nnodes * ndof, 11))
    x = rng.random(gr.nel)
    K = pmb.AssembleGeneral(dom, Ke, bc=bc)(x)
    Ks = K.tocsr()
    n = K.shape[0]
    v = rng.standard_normal(n)
    b = rng.standard_normal(n)
    scale = np.abs(Ks).dot(np.abs(v)).max()
    # rtol 1e-13 of the row magnitude: same products, different summation grouping than scipy's csr_matvec
    np.testing.assert_allclose(K @ v, Ks @ v, rtol=0, atol=1e-13 * scale)
    vd, bd = dv.to_device(v), dv.to_device(b)
    r = dv.empty(n)
    K.apply(_lib.RESIDUAL, vd, r, b=bd)
    np.testing.assert_allclose(r.cpu().numpy(), b - Ks @ v, rtol=0, atol=1e-13 * scale)
    D = K.diagonal_device()
    assert np.array_equal(D.cpu().numpy(), Ks.diagonal())
    u2 = dv.empty(n)
    d3 = dv.empty(3)
    K.apply(_lib.JACOBI, vd, u2, b=bd, diag=D, w=0.5, dotv=bd, dot_out=d3)
    ref = v + 0.5 * ((b - Ks @ v) / Ks.diagonal())
    np.testing.assert_allclose(u2.cpu().numpy(), ref, rtol=0, atol=1e-12 * np.abs(ref).max())
    got = d3.cpu().numpy()
    want = np.array([ref @ v, v @ b, ref @ b])
    np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-9)
    # Dirichlet detection == get_diagonal_indices (solvers.py:88-96)
    lda = pmb.solvers.LDAWrapper(pmb.solvers.Preconditioner())
    lda.update(K)
    assert np.array_equal(lda.diagonal_idx, np.flatnonzero(oracle.solvers.diagonal_only_rows(Ks)))
    assert np.array_equal(lda.diagonal_idx, bc)


def test_transfer_and_galerkin_vs_golden(pmb):
    import torch
    from pymoto_b200 import _lib, device as dv
    from pymoto_b200.matrix import make_grid, DeviceCSR

    g = load("transfer")
    for name in ["2d", "3d", "3d1"]:
        shape = [int(v) for v in g[name + "_shape"]]
        ndof = int(g[name + "_ndof"])
        dom = pmb.VoxelDomain(*shape)
        gf = make_grid(shape[0], shape[1], shape[2], ndof)
        gc = make_grid(shape[0] // 2, shape[1] // 2, shape[2] // 2, ndof)
        v, vc = dv.to_device(g[name + "_v"]), dv.to_device(g[name + "_vc"])
        rc = dv.empty(vc.numel())
        _lib.call("pmb_restrict", gf, gc, dv.ptr(v), dv.ptr(rc), dv.stream())
        assert np.array_equal(rc.cpu().numpy(), g[name + "_restrict"])  # bit-exact (csc_matvec order)
        uf = dv.zeros(v.numel())
        _lib.call("pmb_prolong_add", gf, gc, dv.ptr(vc), dv.ptr(uf), dv.stream())
        assert np.array_equal(uf.cpu().numpy(), g[name + "_prolong"])
        ones = dv.to_device(np.ones(v.numel()))
        _lib.call("pmb_restrict", gf, gc, dv.ptr(ones), dv.ptr(rc), dv.stream())
        assert np.array_equal(rc.cpu().numpy(), g[name + "_restrict_ones"])
        # Galerkin product vs scipy's R^T K R from the reference
        asm = pmb.AssemblePoisson(dom) if ndof == 1 else pmb.AssembleStiffness(dom)
        K = asm(g[name + "_x"])
        mg = pmb.solvers.GeometricMultigrid(dom)
        mg.update(K)
        Ac = mg.Ac.tocsr()
        assert np.array_equal(Ac.indptr, g[name + "_Ac_indptr"]) and np.array_equal(Ac.indices, g[name + "_Ac_indices"])
        ref = g[name + "_Ac_data"]
        # values rtol 1e-12 relative to the largest entry (entries that cancel to ~0 are compared absolutely)
        np.testing.assert_allclose(Ac.data, ref, rtol=1e-12, atol=1e-13 * np.abs(ref).max())


@pytest.mark.gpu
@pytest.mark.parametrize("shape,ndof,nbc", [((6, 4, 4), 3, 11), ((10, 6, 8), 3, 40), ((66, 4, 4), 3, 25), ((8, 8, 8), 1, 9), ((12, 6, 4), 3, 0),
                                            ((70, 6, 6), 1, 30)])
def test_galerkin_direct_vs_scipy(pmb, shape, ndof, nbc):
    """pmb_galerkin_direct (level 1 straight from the element densities, Dirichlet patterns through the table lookup, the
    bc diagonal term through pmb_scatter_add) against scipy's R^T K R with the oracle's prolongation matrix, and against
    the generic two-pass product of the same operator: values rtol 1e-12 (of the largest entry for cancelling ones)."""
    from oracle.solvers import prolongation_matrix
    from pymoto_b200.solvers import GeometricMultigrid

    rng = np.random.default_rng(11)
    gr, gc = Grid(*shape), Grid(*(v // 2 for v in shape))
    dom = pmb.VoxelDomain(*shape)
    Ke = rng.standard_normal((8 * ndof,) * 2)
    Ke = Ke + Ke.T + 8 * np.eye(Ke.shape[0])
    bc = np.unique(rng.integers(0, gr.nnodes * ndof, nbc)) if nbc else None
    if nbc == 40:  # a fully clamped face as well (the cantilever pattern)
        bc = np.unique(np.concatenate([bc, (gr.nodes3d()[0, :, :].ravel()[:, None] * ndof + np.arange(ndof)).ravel()]))
    K = pmb.AssembleGeneral(dom, Ke, bc=bc)(rng.random(gr.nel))
    R = prolongation_matrix(gr, gc, ndof)
    ref = (R.T @ K.tocsr() @ R).tocsr()
    ref.sort_indices()
    assert GeometricMultigrid.direct_level1
    mg = GeometricMultigrid(dom)
    mg.update(K)
    Ad = mg.Ac.tocsr()
    assert np.array_equal(Ad.indptr, ref.indptr) and np.array_equal(Ad.indices, ref.indices)
    np.testing.assert_allclose(Ad.data, ref.data, rtol=1e-12, atol=1e-13 * np.abs(ref.data).max())
    try:
        GeometricMultigrid.direct_level1 = False
        mg2 = GeometricMultigrid(dom)
        mg2.update(K)
        np.testing.assert_allclose(mg2.Ac.data.cpu().numpy(), Ad.data, rtol=1e-12, atol=1e-13 * np.abs(ref.data).max())
    finally:
        GeometricMultigrid.direct_level1 = True


def test_restriction_constants(pmb):
    """reference tests/test_solvers_multigrid.py:9-91 on the kernels: restriction of ones = 8/6/4.5/3.375 (3-D)."""
    from pymoto_b200 import _lib, device as dv
    from pymoto_b200.matrix import make_grid

    gf, gc = make_grid(4, 6, 8, 1), make_grid(2, 3, 4, 1)
    ones = dv.to_device(np.ones(5 * 7 * 9))
    rc = dv.empty(3 * 4 * 5)
    _lib.call("pmb_restrict", gf, gc, dv.ptr(ones), dv.ptr(rc), dv.stream())
    r = rc.cpu().numpy().reshape(5, 4, 3)
    assert np.all(r[1:-1, 1:-1, 1:-1] == 8.0) and r[0, 0, 0] == 3.375
    assert np.all(r[0, 1:-1, 1:-1] == 6.0) and np.all(r[0, 0, 1:-1] == 4.5)
    uf = dv.zeros(5 * 7 * 9)
    ones_c = dv.to_device(np.ones(60))
    _lib.call("pmb_prolong_add", gf, gc, dv.ptr(ones_c), dv.ptr(uf), dv.stream())
    assert np.all(uf.cpu().numpy() == 1.0)


def test_dense_inverse_and_vector_kernels(pmb):
    import torch
    from pymoto_b200 import _lib, device as dv

    rng = np.random.default_rng(3)
    n = 157
    M = rng.standard_normal((n, n))
    M = M @ M.T + n * np.eye(n)
    Md = dv.to_device(M.ravel().copy())
    scratch, info = dv.empty(_lib.query("pmb_dense_invert_ws_doubles", n)), dv.zeros(1, torch.int32)
    _lib.call("pmb_dense_invert", n, dv.ptr(Md), dv.ptr(scratch), dv.ptr(info), dv.stream())
    assert int(info.item()) == 0
    np.testing.assert_allclose(Md.cpu().numpy().reshape(n, n), np.linalg.inv(M), rtol=0, atol=1e-12)
    x = rng.standard_normal(n)
    y = dv.empty(n)
    xd = dv.to_device(x)
    _lib.call("pmb_dense_gemv", n, dv.ptr(Md), dv.ptr(xd), dv.ptr(y), dv.stream())
    np.testing.assert_allclose(y.cpu().numpy(), np.linalg.solve(M, x), rtol=1e-10)
    # dots / lincomb, including empty and ragged lengths
    for m in [1, 31, 1000, 300001]:
        a, b, c = (rng.standard_normal(m) for _ in range(3))
        ad, bd, cd = dv.to_device(a), dv.to_device(b), dv.to_device(c)
        d = dv.dots([(ad, bd), (bd, bd), (ad, cd), (cd, cd)]).cpu().numpy()
        np.testing.assert_allclose(d, [a @ b, b @ b, a @ c, c @ c], rtol=1e-12, atol=1e-12 * m)
        d2 = dv.dots([(ad, bd), (bd, bd)])
        out = dv.empty(m)
        dv.lincomb(out, 2.0, ad, _lib.coef(-1.0, num=dv.scalar_ptr(d2, 0), den=dv.scalar_ptr(d2, 1)), bd)
        np.testing.assert_allclose(out.cpu().numpy(), 2 * a - (a @ b) / (b @ b) * b, rtol=1e-12, atol=1e-12)
        dv.lincomb(out, _lib.coef(1.0, den=dv.scalar_ptr(d2, 1), sqrt_den=True), ad)
        np.testing.assert_allclose(out.cpu().numpy(), a / np.sqrt(b @ b), rtol=1e-13)
    # two identical calls give bit-identical reductions (deterministic partial order)
    a = dv.to_device(rng.standard_normal(1234567))
    assert torch.equal(dv.dots([(a, a)]), dv.dots([(a, a)]))


# ------------------------------------------------------------------------------------------------ filter
@pytest.mark.parametrize("case", CASES)
def test_filter_bit_exact_vs_golden(pmb, case):
    g = load(case)
    nx, ny, nz = (int(v) for v in g["shape"])
    flt = pmb.DensityFilter(pmb.VoxelDomain(nx, ny, nz), radius=float(g["radius"]))
    for dname in DESIGNS:
        assert np.array_equal(flt(g[dname + "_x"]), g[dname + "_y"])
    assert np.array_equal(flt._sensitivity(g["random_filter_bwd_in"]), g["random_filter_bwd"])


@pytest.mark.parametrize("shape,radius", [((37, 5, 3), 2.0), ((9, 7, 0), 1.5), ((4, 3, 2), 3.7), ((33, 34, 9), 2.5), ((3, 3, 0), 1.0)])
def test_filter_vs_oracle_ragged(pmb, shape, radius):
    rng = np.random.default_rng(9)
    gr = Grid(*shape)
    fo = oracle.filter.DensityFilter(gr, radius)
    fg = pmb.DensityFilter(pmb.VoxelDomain(*shape), radius=radius)
    x = rng.random(gr.nel)
    assert np.array_equal(fg.Hs.cpu().numpy(), np.asarray(fo.Hs).ravel())
    assert np.array_equal(fg(x), fo(x))
    assert np.array_equal(fg._sensitivity(x - 0.5), fo.sensitivity(x - 0.5))
    # nonpadding override (filter.py:253-255)
    keep = np.arange(0, gr.nel, 3)
    assert np.array_equal(pmb.DensityFilter(pmb.VoxelDomain(*shape), radius=radius, nonpadding=keep)(x),
                          oracle.filter.DensityFilter(gr, radius, nonpadding=keep)(x))


# ------------------------------------------------------------------------------------------------ solvers
def test_jacobi_cg_vs_golden(pmb):
    """CG(DampedJacobi) against the reference's solution (reference tests/test_solvers_sparse.py:276 style)."""
    g = load("jacobi_cg")
    dom = pmb.VoxelDomain(*[int(v) for v in g["shape"]])
    K = pmb.AssembleStiffness(dom, bc=g["bc"])(g["x"])
    cg = pmb.solvers.CG(K, preconditioner=pmb.solvers.DampedJacobi(K, w=1.0), tol=1e-10)
    u = cg.solve(g["f"])
    np.testing.assert_allclose(u, g["u"], rtol=0, atol=1e-8 * np.abs(g["u"]).max())
    Ks = K.tocsr()
    assert np.linalg.norm(Ks @ u - g["f"]) / np.linalg.norm(g["f"]) <= 1e-10
    with pytest.raises(TypeError):
        cg.solve(g["f"], trans="X")


@pytest.mark.parametrize("case", CASES)
def test_design_iteration_vs_golden(pmb, case):
    """Full chain through the Module API: filter -> SIMP -> assembly -> LinSolve(CG(GMG)) -> compliance -> sensitivities."""
    g = load(case)
    nx, ny, nz = (int(v) for v in g["shape"])
    kind, xmin, tol, min_size = str(g["kind"]), float(g["xmin"]), float(g["tol"]), int(g["min_size"])
    dom = pmb.VoxelDomain(nx, ny, nz)
    ndof = int(g["ndof"])
    f = force_vector(g, dom.nnodes * ndof)
    P = ComplianceProblem(Grid(nx, ny, nz), kind=kind, radius=float(g["radius"]), xmin=xmin, tol=tol, min_size=min_size)
    for dname in DESIGNS:
        x = g[dname + "_x"]
        sx = pmb.Signal("x", state=x.copy())
        with pmb.Network() as fn:
            sy = pmb.DensityFilter(dom, radius=float(g["radius"]))(sx)
            ss = _Simp(xmin)(sy)
            asm = pmb.AssemblePoisson(dom, bc=g["bc"]) if kind == "heatsink" else pmb.AssembleStiffness(dom, bc=g["bc"])
            sK = asm(ss)
            mgs = pmb.solvers.auto_multigrid(dom, min_size=min_size)
            assert len(mgs) == int(g["n_mg"])
            cg = pmb.solvers.CG(preconditioner=mgs[0], tol=tol)
            su = pmb.LinSolve(hermitian=True, solver=cg)(sK, f)
            sc = _Dot()(su, f)
        u = su.state
        Ks = sK.state.tocsr()
        relres = np.linalg.norm(Ks @ u - f) / np.linalg.norm(f)
        assert relres <= 1e-8
        c = float(sc.state)
        assert abs(c - float(g[dname + "_compliance"])) <= 1e-6 * abs(float(g[dname + "_compliance"]))
        np.testing.assert_allclose(u, g[dname + "_u"], rtol=0, atol=1e-6 * np.abs(g[dname + "_u"]).max())
        # same iteration as the reference: CG iteration count within +-1 of the CPU oracle on the same input
        P.u = None
        P.response(x)
        assert abs(cg.iterations - P.cg.iterations) <= 1, (cg.iterations, P.cg.iterations)
        # backward
        sc.sensitivity = 1.0
        fn.sensitivity()
        assert not su_solver(fn)._did_solve  # adjoint from the LDAS database: no CG
        ref = g[dname + "_dcdx"]
        np.testing.assert_allclose(sx.sensitivity, ref, rtol=1e-6, atol=1e-7 * np.abs(ref).max())
        # a second response with the unchanged design converges in 0 iterations from the warm start
        fn.reset()
        fn.response()
        assert cg.iterations == 0


def su_solver(fn):
    for m in fn.mods:
        if type(m).__name__ == "LinSolve":
            return m.solver
    raise AssertionError


def _make_glue():
    import pymoto_b200 as pmb

    class Simp(pmb.Module):
        """Host glue standing in for pym.MathExpression("xmin + (1-xmin)*inp0^3") (generic.py:94-140)."""

        def __init__(self, xmin):
            self.xmin = xmin

        def __call__(self, y):
            self.y = y
            return self.xmin + (1.0 - self.xmin) * y ** 3

        def _sensitivity(self, ds):
            return ds * (3.0 * (1.0 - self.xmin) * self.y ** 2)

    class Dot(pmb.Module):
        """pym.EinSum('i,i->') (generic.py:143-226)."""

        def __call__(self, a, b):
            self.a, self.b = a, b
            return a @ b

        def _sensitivity(self, dc):
            return dc * self.b, dc * self.a

    return Simp, Dot


def _Simp(xmin):
    return _make_glue()[0](xmin)


def _Dot():
    return _make_glue()[1]()


def test_sensitivity_kernel_vs_oracle(pmb):
    from pymoto_b200 import device as dv

    rng = np.random.default_rng(17)
    for shape, ndof in [((5, 4, 3), 3), ((9, 6, 0), 2), ((4, 4, 4), 1)]:
        gr = Grid(*shape)
        dom = pmb.VoxelDomain(*shape)
        Ke = rng.standard_normal((gr.elemnodes * ndof,) * 2)
        bc = np.unique(rng.integers(0, gr.nnodes * ndof, 9))
        asm = pmb.AssembleGeneral(dom, Ke, bc=bc)
        asm(rng.random(gr.nel))
        u, v = rng.standard_normal(gr.nnodes * ndof), rng.standard_normal(gr.nnodes * ndof)
        u2, v2 = rng.standard_normal(gr.nnodes * ndof), rng.standard_normal(gr.nnodes * ndof)
        dyad = pmb.DeviceDyad([dv.to_device(u), dv.to_device(u2)], [dv.to_device(v), dv.to_device(v2)])
        dx = asm._sensitivity(dyad)[0]
        oa = oracle.assembly.Assembler(gr, Ke, bc=bc)
        ref = oa.sensitivity(u, v) + oa.sensitivity(u2, v2)
        np.testing.assert_allclose(dx, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())


def test_golden_compliance_64x32x32(pmb):
    """64x32x32 cantilever, x = 0.5, tol 1e-8 -> compliance 10528.606127394825 (reference, 3 MG levels)."""
    import torch
    from pymoto_b200 import device as dv

    nx, ny, nz = 64, 32, 32
    dom = pmb.VoxelDomain(nx, ny, nz)
    ndof, bc, f = cantilever(Grid(nx, ny, nz))
    flt = pmb.DensityFilter(dom, radius=2.0)
    asm = pmb.AssembleStiffness(dom, bc=bc)
    mgs = pmb.solvers.auto_multigrid(dom)
    assert len(mgs) == 3
    cg = pmb.solvers.CG(preconditioner=mgs[0], tol=1e-8)
    ls = pmb.LinSolve(hermitian=True, solver=cg)
    x = dv.to_device(np.full(dom.nel, 0.5))
    y = flt(x)
    s = 1e-9 + (1.0 - 1e-9) * y ** 3
    K = asm(s)
    fd = dv.to_device(f)
    u = ls(K, fd)
    c = float(u @ fd)
    assert abs(c - 10528.606127394825) <= 1e-6 * 10528.606127394825
    assert pmb.solvers.LinearSolver.residual(K, u, fd) <= 1e-8
    assert 6 <= cg.iterations <= 9  # reference: 7 products ("6 iterations" printed 0-based)
    # size-independent properties at this size: symmetry of the operator and linearity of the filter
    rng = np.random.default_rng(0)
    a, b = dv.to_device(rng.standard_normal(K.shape[0])), dv.to_device(rng.standard_normal(K.shape[0]))
    lhs, rhs = float((K @ a) @ b), float(a @ (K @ b))
    assert abs(lhs - rhs) <= 1e-10 * max(abs(lhs), 1.0)
    x2 = dv.to_device(rng.random(dom.nel))
    assert torch.allclose(flt(x + x2), flt(x) + flt(x2), rtol=1e-13, atol=1e-13)


def test_dropin_under_real_pymoto_runtime():
    """INTEGRATION.md section 3 verbatim under the unmodified reference's own Module / Network runtime and MMA optimiser
    (fresh interpreter: pymoto must be importable BEFORE pymoto_b200 so that the hot-path classes derive from pymoto.Module)."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "baseline"))
    import refload

    if refload.reference_root() is None:
        pytest.skip("reference not installed (python baseline/install_ref.py)")
    res = subprocess.run([sys.executable, os.path.join(root, "tests", "dropin_check.py")], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "[dropin_check] OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


def test_pure_c_abi_solve_without_torch():
    """A host with no torch in the process (cudaMalloc through ctypes) builds the multigrid hierarchy and solves 32x16x16 with
    pmb_pcg_solve through the C ABI alone; iterations, residual and solution against the oracle."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tests", "c_abi_check.py")], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "[c_abi_check] OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


def test_c_abi_comm_entry_points_two_gpus(tmp_path):
    """pmb_comm_init / pmb_halo_exchange / pmb_allreduce (NCCL bound at run time) from two plain processes, one per GPU."""
    import subprocess
    import sys
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    idfile = str(tmp_path / "nccl_id.bin")
    procs = [subprocess.Popen([sys.executable, os.path.join(root, "tests", "c_abi_check.py"), "comm", str(r), "2", idfile],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"comm rank {r}/2 OK" in o, o[-3000:]


def test_multi_gpu_slab_parity():
    """z-slab path on 2 GPUs vs the oracle (skipped on a single-GPU box; run with `gpurun --gpus 2`)."""
    import os
    import subprocess
    import sys
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(root, "tests", "dist_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and "[dist_check] OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


@pytest.mark.parametrize("shape,ndof", [((6, 4, 4), 3), ((12, 8, 0), 2), ((8, 8, 8), 1), ((33, 9, 5), 3), ((70, 11, 0), 1),
                                        ((40, 6, 3), 1), ((35, 18, 3), 2)])
def test_matrix_free_operator_equals_assembled(pmb, shape, ndof):
    """pmb_elem_spmv (finest level evaluated from the element scaling) against the assembled CSR operator, all modes
    and the fused dot products; rtol 1e-13 of the row magnitude (different summation grouping, same products)."""
    from pymoto_b200 import _lib, device as dv
    from pymoto_b200.matrix import DeviceCSR

    rng = np.random.default_rng(2)
    gr = Grid(*shape)
    dom = pmb.VoxelDomain(*shape)
    Ke = rng.standard_normal((gr.elemnodes * ndof,) * 2)
    Ke = Ke + Ke.T + 8 * np.eye(Ke.shape[0])
    bc = np.unique(rng.integers(0, gr.nnodes * ndof, 13))
    K = pmb.AssembleGeneral(dom, Ke, bc=bc)(rng.random(gr.nel))
    assert K.generator is not None
    Ks = K.tocsr()
    n = K.shape[0]
    v, b = rng.standard_normal(n), rng.standard_normal(n)
    vd, bd = dv.to_device(v), dv.to_device(b)
    D = K.diagonal_device()
    scale = np.abs(Ks).dot(np.abs(v)).max()
    for mode, ref in [(_lib.SPMV, Ks @ v), (_lib.RESIDUAL, b - Ks @ v), (_lib.JACOBI, v + 0.5 * ((b - Ks @ v) / Ks.diagonal()))]:
        out, d3 = dv.empty(n), dv.empty(3)
        assert DeviceCSR.matrix_free
        K.apply(mode, vd, out, b=bd, diag=D, w=0.5, dotv=bd, dot_out=d3)
        np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=0, atol=2e-13 * max(scale, np.abs(ref).max()))
        np.testing.assert_allclose(d3.cpu().numpy(), [ref @ v, v @ b, ref @ b], rtol=1e-10, atol=1e-9)
        try:  # and the CSR-streaming kernel on the same matrix
            DeviceCSR.matrix_free = False
            out2 = dv.empty(n)
            K.apply(mode, vd, out2, b=bd, diag=D, w=0.5)
        finally:
            DeviceCSR.matrix_free = True
        np.testing.assert_allclose(out2.cpu().numpy(), out.cpu().numpy(), rtol=0, atol=2e-13 * max(scale, np.abs(ref).max()))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,ndof", [((37, 9, 7), 3), ((33, 6, 4), 3), ((40, 7, 6), 1), ((64, 32, 32), 3), ((5, 20, 19), 3), ((34, 9, 17), 2),
                                        ((66, 16, 5), 3), ((31, 8, 4), 3)])
def test_matrix_free_kernel_variants_bit_identical(pmb, shape, ndof):
    """Every layout of the 3-D matrix-free kernel must reproduce variant 0 (one node per thread on a brick): the z-marching
    columns (1, 2) and the bulk-copy ring layouts (4, 5; with and without brick flags) bit for bit, the FP64 tensor-core layout (3; DMMA accumulation order) to 1e-11 of the field magnitude in y (same products, same order per node) for all modes, on the whole grid and
    on sub-slabs with odd plane counts; the fused dot products agree to rounding.  The autotune entry point runs, returns
    one time per variant and leaves a valid selection."""
    import ctypes as C

    import torch
    from pymoto_b200 import _lib, device as dv
    from pymoto_b200.matrix import make_grid

    rng = np.random.default_rng(5)
    gr = Grid(*shape)
    dom = pmb.VoxelDomain(*shape)
    Ke = rng.standard_normal((8 * ndof,) * 2)
    Ke = Ke + Ke.T + 8 * np.eye(Ke.shape[0])
    bc = np.unique(rng.integers(0, gr.nnodes * ndof, 17))
    K = pmb.AssembleGeneral(dom, Ke, bc=bc)(rng.random(gr.nel))
    n = K.shape[0]
    vd, bd = dv.to_device(rng.standard_normal(n)), dv.to_device(rng.standard_normal(n))
    D = K.diagonal_device()
    gen = K.generator
    saved = gen.variant
    try:
        ref = {}
        nvar = _lib.query("pmb_elem_num_variants")
        assert nvar >= 8
        for variant in range(nvar):
            gen.variant = variant
            for mode in (_lib.SPMV, _lib.RESIDUAL, _lib.JACOBI):
                out, d3 = dv.zeros(n), dv.empty(3)
                K.apply(mode, vd, out, b=bd, diag=D, w=0.5, dotv=bd, dot_out=d3)
                got = (out.cpu().numpy(), d3.cpu().numpy())
                exact = variant not in (3, 6, 7) or ndof != 3
                if variant == 0:
                    ref[mode] = got
                elif exact:
                    assert np.array_equal(got[0], ref[mode][0]), (variant, mode)
                    np.testing.assert_allclose(got[1], ref[mode][1], rtol=1e-11, atol=1e-9)
                else:
                    np.testing.assert_allclose(got[0], ref[mode][0], rtol=0, atol=1e-11 * max(1.0, np.abs(ref[mode][0]).max()))
                    np.testing.assert_allclose(got[1], ref[mode][1], rtol=1e-10, atol=1e-8)
            # a sub-slab [k0, k0 + 3) of the same operator: pointers move with the slab, halo planes are read around it
            g, k0, npl = K.grid, 1, 3
            sg = make_grid(g.nx, g.ny, g.nz, g.ndof, k0, npl)
            plane = K.plane
            out = dv.zeros(n)
            xin = K._padded(vd)
            _lib.call("pmb_elem_spmv", sg, _lib.JACOBI, C.byref(gen.op(sg, k0)), xin.data_ptr() + 8 * k0 * plane,
                      bd.data_ptr() + 8 * k0 * plane, D.data_ptr() + 8 * k0 * plane, 0.5, out.data_ptr() + 8 * k0 * plane,
                      None, None, None, dv.stream())
            got = out.cpu().numpy()
            want = np.zeros(n)
            want[k0 * plane:(k0 + npl) * plane] = ref[_lib.JACOBI][0][k0 * plane:(k0 + npl) * plane]
            if exact:
                assert np.array_equal(got, want), ("slab", variant)
            else:
                np.testing.assert_allclose(got, want, rtol=0, atol=1e-11 * max(1.0, np.abs(want).max()))
            if variant in (4, 5, 6, 7):  # without brick flags every brick applies the mask: same result
                op = gen.op()
                bare = _lib.ElemOp(op.Ke_host, op.s, op.bcmask, op.bcdiagval, None, variant)
                out = dv.zeros(n)
                _lib.call("pmb_elem_spmv", K.grid, _lib.JACOBI, C.byref(bare), xin.data_ptr(), bd.data_ptr(), D.data_ptr(), 0.5,
                          out.data_ptr(), None, None, None, dv.stream())
                if exact:
                    assert np.array_equal(out.cpu().numpy(), ref[_lib.JACOBI][0]), ("no flags", variant)
                else:
                    np.testing.assert_allclose(out.cpu().numpy(), ref[_lib.JACOBI][0], rtol=0, atol=1e-11 * max(1.0, np.abs(ref[_lib.JACOBI][0]).max()))
        ms = (C.c_double * nvar)()
        best = C.c_int(-1)
        scratch = dv.zeros(n)
        fscr = dv.empty(_lib.query("pmb_elem_autotune_flag_bytes", K.grid), torch.uint8)
        for allow in (1, 0):
            _lib.call("pmb_elem_autotune", K.grid, C.byref(gen.op()), xin.data_ptr(), bd.data_ptr(), D.data_ptr(), scratch.data_ptr(),
                      fscr.data_ptr(), allow, C.addressof(ms), C.byref(best), dv.stream())
            assert all(0.0 < t < 1e3 for t in ms)
            assert 0 <= best.value < nvar
        assert best.value not in (3, 6, 7) or ndof != 3
        np.testing.assert_allclose(scratch.cpu().numpy(), ref[_lib.JACOBI][0], rtol=0, atol=1e-11 * max(1.0, np.abs(ref[_lib.JACOBI][0]).max()))
    finally:
        gen.variant = saved


@pytest.mark.gpu
@pytest.mark.parametrize("shape,ndof,units", [((37, 9, 7), 3, (1.0, 1.0, 1.0)), ((33, 6, 4), 3, (1.0, 0.5, 2.0)), ((40, 7, 6), 1, (1.0, 1.0, 1.0)),
                                              ((64, 32, 32), 3, (1.0, 1.0, 1.0)), ((5, 20, 19), 3, (0.25, 1.0, 1.5)), ((66, 16, 5), 1, (2.0, 0.5, 1.0)),
                                              ((31, 8, 4), 3, (1.0, 1.0, 1.0)), ((30, 7, 33), 3, (1.0, 1.0, 1.0)), ((62, 14, 3), 3, (1.0, 1.0, 1.0))])
def test_matrix_free_parity_block_layout(pmb, shape, ndof, units):
    """Layouts 8 / 9 (parity-block form of the element product, pmb_elem_par.cuh) on the reference's own hex8 stiffness /
    conductivity matrices (cubic and cuboid voxels): every mode, the fused dot products and a sub-slab agree with layout 0 to
    1e-12 of the field magnitude; an element matrix WITHOUT the reflection symmetry silently runs layout 0 (bit-identical)."""
    import ctypes as C

    from pymoto_b200 import _lib, device as dv
    from pymoto_b200.matrix import make_grid

    rng = np.random.default_rng(11)
    gr = Grid(*shape)
    dom = pmb.VoxelDomain(*shape, unitx=units[0], unity=units[1], unitz=units[2])
    bc = np.unique(rng.integers(0, gr.nnodes * ndof, 23))
    asm = (pmb.AssembleStiffness if ndof == 3 else pmb.AssemblePoisson)(dom, bc=bc)
    K = asm(rng.random(gr.nel) * 0.9 + 0.1)
    n = K.shape[0]
    vd, bd = dv.to_device(rng.standard_normal(n)), dv.to_device(rng.standard_normal(n))
    D = K.diagonal_device()
    gen = K.generator
    saved = gen.variant
    try:
        ref = {}
        for variant in (0, 8, 9):
            gen.variant = variant
            for mode in (_lib.SPMV, _lib.RESIDUAL, _lib.JACOBI):
                out, d3 = dv.zeros(n), dv.empty(3)
                K.apply(mode, vd, out, b=bd, diag=D, w=0.5, dotv=bd, dot_out=d3)
                got = (out.cpu().numpy(), d3.cpu().numpy())
                if variant == 0:
                    ref[mode] = got
                else:
                    scale = max(1.0, np.abs(ref[mode][0]).max())
                    np.testing.assert_allclose(got[0], ref[mode][0], rtol=0, atol=1e-12 * scale, err_msg=f"variant {variant} mode {mode}")
                    assert np.abs(got[0] - ref[mode][0]).max() > 0.0 or n < 100, "layout 0 ran instead of the parity-block layout"
                    np.testing.assert_allclose(got[1], ref[mode][1], rtol=1e-10, atol=1e-8)
            if variant == 0 or K.grid.nz < 5:
                continue
            g, k0, npl = K.grid, 1, 3
            sg = make_grid(g.nx, g.ny, g.nz, g.ndof, k0, npl)
            plane = K.plane
            out = dv.zeros(n)
            xin = K._padded(vd)
            _lib.call("pmb_elem_spmv", sg, _lib.JACOBI, C.byref(gen.op(sg, k0)), xin.data_ptr() + 8 * k0 * plane,
                      bd.data_ptr() + 8 * k0 * plane, D.data_ptr() + 8 * k0 * plane, 0.5, out.data_ptr() + 8 * k0 * plane,
                      None, None, None, dv.stream())
            want = np.zeros(n)
            want[k0 * plane:(k0 + npl) * plane] = ref[_lib.JACOBI][0][k0 * plane:(k0 + npl) * plane]
            np.testing.assert_allclose(out.cpu().numpy(), want, rtol=0, atol=1e-12 * max(1.0, np.abs(want).max()))
        # no reflection symmetry (x displacement coupled to a y difference with one sign only): layout 0 runs
        Ke = np.array(asm._Ke_host, copy=True).reshape(8 * ndof, 8 * ndof)
        Ke[0, -1] += 0.05
        Ke[-1, 0] += 0.05
        K2 = pmb.AssembleGeneral(dom, Ke, bc=bc)(rng.random(gr.nel))
        outs = []
        for variant in (0, 8):
            K2.generator.variant = variant
            out = dv.zeros(n)
            K2.apply(_lib.SPMV, vd, out)
            outs.append(out.cpu().numpy())
        assert np.array_equal(outs[0], outs[1])
    finally:
        gen.variant = saved


@pytest.mark.gpu
@pytest.mark.parametrize("shape,ndof", [((6, 4, 4), 3), ((8, 8, 8), 1), ((33, 9, 5), 3), ((130, 5, 3), 1), ((20, 18, 3), 2), ((37, 11, 9), 3)])
def test_symmetric_half_stencil_storage(pmb, shape, ndof):
    """pmb_sym_pack / pmb_sym_spmv (symmetric half-stencil layout of the coarse-level operators): every mode and the fused dot
    products against scipy on the assembled CSR and against the stencil-CSR kernel; the measured asymmetry is ~0 for a
    symmetric operator and the layout is refused (stencil-CSR kernel keeps running) for a non-symmetric one."""
    from pymoto_b200 import _lib, device as dv
    from pymoto_b200.matrix import DeviceCSR

    rng = np.random.default_rng(3)
    gr = Grid(*shape)
    dom = pmb.VoxelDomain(*shape)
    Ke = rng.standard_normal((gr.elemnodes * ndof,) * 2)
    Ke = Ke + Ke.T + 8 * np.eye(Ke.shape[0])
    bc = np.unique(rng.integers(0, gr.nnodes * ndof, 11))
    K0 = pmb.AssembleGeneral(dom, Ke, bc=bc)(rng.random(gr.nel))
    Ks = K0.tocsr()
    n = K0.shape[0]
    saved = DeviceCSR.symmetric_min_nodes, DeviceCSR.symmetric_storage
    DeviceCSR.symmetric_min_nodes, DeviceCSR.symmetric_storage = 0, True
    try:
        K = DeviceCSR(K0.grid, data=K0._buf, level=1)  # the same values seen as a coarse-level operator
        assert K.pack_symmetric() and K.asymmetry <= 1e-15
        v, b = rng.standard_normal(n), rng.standard_normal(n)
        vd, bd = dv.to_device(v), dv.to_device(b)
        scale = np.abs(Ks).dot(np.abs(v)).max()
        D = K.diagonal_device()
        calls0 = sum(c for (nm, _), c in _lib.call_stats.items() if nm == "pmb_sym_spmv")
        for mode, want in ((_lib.SPMV, Ks @ v), (_lib.RESIDUAL, b - Ks @ v), (_lib.JACOBI, v + 0.5 * ((b - Ks @ v) / Ks.diagonal()))):
            out, d3 = dv.zeros(n), dv.empty(3)
            K.apply(mode, vd, out, b=bd, diag=D, w=0.5, dotv=bd, dot_out=d3)
            np.testing.assert_allclose(out.cpu().numpy(), want, rtol=0, atol=1e-13 * scale)
            np.testing.assert_allclose(d3.cpu().numpy(), [want @ v, v @ b, want @ b], rtol=1e-11, atol=1e-9)
            out2 = dv.zeros(n)
            K.apply(mode, vd, out2, b=bd, diag=D, w=0.5)
            assert np.array_equal(out2.cpu().numpy(), out.cpu().numpy())
            ref = dv.zeros(n)
            was, DeviceCSR.matrix_free = DeviceCSR.matrix_free, False  # level 0 through the stencil-CSR kernel
            try:
                K0.apply(mode, vd, ref, b=bd, diag=D, w=0.5)
            finally:
                DeviceCSR.matrix_free = was
            np.testing.assert_allclose(out.cpu().numpy(), ref.cpu().numpy(), rtol=0, atol=2e-13 * scale)
        assert sum(c for (nm, _), c in _lib.call_stats.items() if nm == "pmb_sym_spmv") == calls0 + 6
        # a non-symmetric operator: the layout is refused and the stencil-CSR kernel runs
        data = K0.data.clone()
        data[::7] *= 1.5
        buf = dv.empty(K0.nnz + 2)
        buf[:K0.nnz], buf[K0.nnz:] = data, 0.0
        Kn = DeviceCSR(K0.grid, data=buf, level=1)
        assert not Kn.pack_symmetric() and Kn.asymmetry > 1e-3
        out = dv.zeros(n)
        Kn.apply(_lib.SPMV, vd, out)
        import scipy.sparse as sp

        Kn_s = sp.csr_matrix((data.cpu().numpy(), Ks.indices, Ks.indptr), shape=Ks.shape)
        np.testing.assert_allclose(out.cpu().numpy(), Kn_s @ v, rtol=0, atol=2e-13 * 1.5 * scale)
    finally:
        DeviceCSR.symmetric_min_nodes, DeviceCSR.symmetric_storage = saved


@pytest.mark.gpu
def test_multigrid_solve_with_symmetric_storage_matches(pmb):
    """The CG + GMG solve with the coarse levels swept from the symmetric half-stencil copy: same iteration count, same
    solution (1e-9) and compliance (1e-12) as with the stencil-CSR sweeps."""
    from pymoto_b200 import _lib
    from pymoto_b200.matrix import DeviceCSR

    nx, ny, nz = 32, 16, 16
    dom = pmb.VoxelDomain(nx, ny, nz)
    ndof, bc, f = cantilever(Grid(nx, ny, nz))
    rng = np.random.default_rng(2)
    xs = 0.2 + 0.8 * rng.random(dom.nel)
    saved = DeviceCSR.symmetric_min_nodes, DeviceCSR.symmetric_storage
    res = {}
    try:
        DeviceCSR.symmetric_min_nodes = 0
        for on in (False, True):
            DeviceCSR.symmetric_storage = on
            K = pmb.AssembleStiffness(dom, bc=bc)(xs ** 3)
            cg = pmb.solvers.CG(preconditioner=pmb.solvers.auto_multigrid(dom)[0], tol=1e-9)
            n0 = sum(c for (nm, _), c in _lib.call_stats.items() if nm == "pmb_sym_spmv")
            u = pmb.LinSolve(hermitian=True, solver=cg)(K, f)
            used = sum(c for (nm, _), c in _lib.call_stats.items() if nm == "pmb_sym_spmv") - n0
            assert (used > 0) == on
            res[on] = (np.asarray(u), cg.iterations)
        np.testing.assert_allclose(res[True][0], res[False][0], rtol=0, atol=1e-9 * np.abs(res[False][0]).max())
        assert abs(f @ res[True][0] - f @ res[False][0]) <= 1e-12 * abs(f @ res[False][0])
        assert res[True][1] == res[False][1]
    finally:
        DeviceCSR.symmetric_min_nodes, DeviceCSR.symmetric_storage = saved


# ------------------------------------------------------------------------------------------------ FilterConv (next row f1)
FILTERCONV_KW = {
    "sym3d": dict(radius=2.0),
    "r3_3d": dict(radius=3.2),
    "sym2d": dict(radius=2.5),
    "edge_wrap": dict(radius=2.0, xmin_bc="edge", xmax_bc="wrap", ymin_bc="wrap", ymax_bc="wrap", zmin_bc="edge", zmax_bc="edge"),
    "const": dict(radius=2.0, xmin_bc=0.0, xmax_bc=1.0, ymin_bc=0.25, ymax_bc="symmetric", zmin_bc="edge", zmax_bc=0.75),
    "weights": dict(xmin_bc="wrap", xmax_bc="symmetric", ymin_bc=0.5, ymax_bc="edge"),
    "weights2d": dict(xmin_bc="edge", ymax_bc=2.0),
    "override": dict(radius=2.0),
}
FILTERCONV_SHAPES = {"sym3d": (7, 5, 4), "r3_3d": (9, 6, 5), "sym2d": (12, 9, 0), "edge_wrap": (8, 6, 5), "const": (6, 7, 5),
                     "weights": (6, 5, 4), "weights2d": (9, 8, 0), "override": (6, 6, 4)}


@pytest.mark.parametrize("name", list(FILTERCONV_KW))
def test_filterconv_vs_reference_golden(pmb, name):
    """Padded convolution filter, every boundary mode, custom (asymmetric) kernels and value overrides, against the
    reference's own outputs (scipy.signal convolve / correlate): rtol 1e-12 of the field magnitude."""
    g = load("filterconv")
    kw = dict(FILTERCONV_KW[name])
    if name + "_weights" in g.files:
        kw["weights"] = g[name + "_weights"]
    m = pmb.FilterConv(pmb.VoxelDomain(*FILTERCONV_SHAPES[name]), **kw)
    if name == "override":
        m.override_values((np.s_[1:3], np.s_[2:4], np.s_[:]), 1.0)
    y = m(g[name + "_x"])
    np.testing.assert_allclose(y, g[name + "_y"], rtol=0, atol=1e-12 * np.abs(g[name + "_y"]).max())
    dx = m._sensitivity(g[name + "_dy"])
    np.testing.assert_allclose(dx, g[name + "_dx"], rtol=0, atol=1e-12 * np.abs(g[name + "_dx"]).max())
    if "radius" in kw and name != "override" and not any(isinstance(v, float) for v in kw.values() if v is not kw["radius"]):
        # volume preserving with symmetric / edge / wrap padding: a constant field is reproduced
        one = m(np.ones(m.nel))
        np.testing.assert_allclose(one, 1.0, rtol=1e-13)
    with pytest.raises(ValueError):
        pmb.FilterConv(pmb.VoxelDomain(4, 4, 4))


# ------------------------------------------------------------------------------------------------ BASELINE configs
def test_config0_mbb_2d_100x50(pmb):
    """BASELINE.json configs[0]: 2-D MBB beam 100x50 quad4, x = 0.5, DensityFilter r=2, SIMP p=3.  The reference (scipy
    direct solve) gives compliance 369.30859599959234; here through CG + one multigrid level."""
    nx, ny = 100, 50
    dom = pmb.VoxelDomain(nx, ny)
    nodes = dom.nodes
    bc = np.concatenate([2 * nodes[0, :].flatten(), 2 * nodes[nx, 0].flatten() + 1])  # u_x = 0 on i = 0, u_y = 0 at (nx, 0)
    f = np.zeros(dom.nnodes * 2)
    f[2 * nodes[0, ny].flatten() + 1] = -1.0
    x = np.full(dom.nel, 0.5)
    y = pmb.DensityFilter(dom, radius=2.0)(x)
    s = 1e-9 + (1.0 - 1e-9) * y ** 3
    K = pmb.AssembleStiffness(dom, bc=bc)(s)
    mgs = pmb.solvers.auto_multigrid(dom)
    cg = pmb.solvers.CG(preconditioner=mgs[0], tol=1e-10)
    u = pmb.LinSolve(hermitian=True, solver=cg)(K, f)
    c = float(u @ f)
    assert abs(c - 369.30859599959234) <= 1e-6 * 369.30859599959234
    Ks = K.tocsr()
    assert np.linalg.norm(Ks @ u - f) / np.linalg.norm(f) <= 1e-8
    # the same system solved directly by the oracle's sparse LU
    import scipy.sparse.linalg as spla

    uref = spla.spsolve(Ks.tocsc(), f)
    np.testing.assert_allclose(u, uref, rtol=0, atol=1e-7 * np.abs(uref).max())


def test_config4_thermal_chain_vs_oracle(pmb):
    """BASELINE.json configs[4] (heat-sink, scalar conduction, AssemblePoisson) at 32^3 against the CPU oracle."""
    nx = ny = nz = 32
    P = ComplianceProblem(Grid(nx, ny, nz), kind="heatsink", tol=1e-8, min_size=8)
    x = np.random.default_rng(8).random(P.grid.nel) * 0.8 + 0.2
    c_ref = P.response(x)
    dx_ref = P.sensitivity()
    dom = pmb.VoxelDomain(nx, ny, nz)
    flt = pmb.DensityFilter(dom, radius=2.0)
    y = flt(x)
    assert np.array_equal(y, P.y)
    s = 1e-9 + (1.0 - 1e-9) * y ** 3
    asm = pmb.AssemblePoisson(dom, bc=P.bc)
    K = asm(s)
    assert np.array_equal(K.data.cpu().numpy(), P.K.data)
    mgs = pmb.solvers.auto_multigrid(dom, min_size=8)
    assert len(mgs) == len(P.mgs)
    cg = pmb.solvers.CG(preconditioner=mgs[0], tol=1e-8)
    ls = pmb.LinSolve(hermitian=True, solver=cg)
    u = ls(K, P.f)
    c = float(u @ P.f)
    assert abs(c - c_ref) <= 1e-6 * abs(c_ref)
    assert abs(cg.iterations - P.cg.iterations) <= 1
    dmat, _ = ls._sensitivity(P.f)
    ds = asm._sensitivity(dmat)[0]
    dx = flt._sensitivity(ds * (3.0 * (1.0 - 1e-9) * y ** 2))
    np.testing.assert_allclose(dx, dx_ref, rtol=1e-6, atol=1e-7 * np.abs(dx_ref).max())


def test_full_size_properties_256x128x128(pmb):
    """Size-independent properties at BASELINE's full size (12.8 M dof, too big for the CPU oracle):
    rigid-body translations lie in the null space of the unconstrained operator, the operator is symmetric, a constant
    field passes the filter unchanged, the matrix-free and the assembled operator agree, and the solve reaches 1e-8."""
    import torch
    from pymoto_b200 import _lib, device as dv
    from pymoto_b200.matrix import DeviceCSR

    nx, ny, nz = 256, 128, 128
    dom = pmb.VoxelDomain(nx, ny, nz)
    rng = np.random.default_rng(0)
    xe = dv.to_device(rng.random(dom.nel) * 0.9 + 0.1)
    flt = pmb.DensityFilter(dom, radius=2.0)
    ones = dv.to_device(np.ones(dom.nel))
    assert torch.allclose(flt(ones), ones, rtol=1e-14, atol=0)  # (H 1)/Hs = 1
    # linearity of the filter transpose pair: <H x, y> = <x, H^T y>
    a, b = flt(xe), flt._sensitivity(xe)
    assert abs(float(a @ xe) - float(xe @ b)) <= 1e-10 * abs(float(a @ xe))
    # unconstrained stiffness: K * (rigid translation) = 0, row sums of the values vanish
    K0 = pmb.AssembleStiffness(dom)(xe)
    n = K0.shape[0]
    t = dv.zeros(n)
    t[2::3] = 1.0
    for mf in (True, False):
        DeviceCSR.matrix_free = mf
        try:
            r = K0 @ t
        finally:
            DeviceCSR.matrix_free = True
        assert float(r.abs().max()) <= 1e-12 * float(K0.diagonal_device().abs().max())
    # checksum of checksums: sum of all assembled values == sum_e s_e * sum(Ke) (= 0 for a stiffness matrix) and
    # trace == sum_e s_e trace(Ke)
    Ke = pmb.AssembleStiffness(dom).elmat[0]
    tr = float(K0.diagonal_device().sum())
    assert abs(tr - float(xe.sum()) * np.trace(Ke)) <= 1e-10 * abs(tr)
    # cantilever solve at full size: residual of the ASSEMBLED matrix <= 1e-8 although CG iterates matrix-free
    ndof, bc, f = cantilever(Grid(nx, ny, nz))
    asm = pmb.AssembleStiffness(dom, bc=bc)
    s = 1e-9 + (1.0 - 1e-9) * flt(xe) ** 3
    K = asm(s)
    mgs = pmb.solvers.auto_multigrid(dom)
    assert len(mgs) == 5
    cg = pmb.solvers.CG(preconditioner=mgs[0], tol=1e-8)
    fd = dv.to_device(f)
    u = pmb.LinSolve(hermitian=True, solver=cg)(K, fd)
    DeviceCSR.matrix_free = False
    try:
        relres = pmb.solvers.LinearSolver.residual(K, u, fd)
    finally:
        DeviceCSR.matrix_free = True
    assert relres <= 1e-8
    # symmetry of the constrained operator
    p, q = dv.to_device(rng.standard_normal(n)), dv.to_device(rng.standard_normal(n))
    lhs, rhs = float((K @ p) @ q), float(p @ (K @ q))
    assert abs(lhs - rhs) <= 1e-9 * max(abs(lhs), abs(rhs), 1.0)


def test_host_numpy_module_chain_network(pmb):
    """The drop-in use: numpy arrays at every module edge inside a Network (the way the reference's scripts run), with
    the reference-style glue modules; response, sensitivity, reset and a second iteration with a changed design."""
    nx, ny, nz = 16, 8, 8
    P = ComplianceProblem(Grid(nx, ny, nz), kind="cantilever", tol=1e-8, min_size=4)
    dom = pmb.VoxelDomain(nx, ny, nz)
    rng = np.random.default_rng(21)
    x0 = rng.random(dom.nel)
    sx = pmb.Signal("x", state=x0.copy())
    with pmb.Network() as fn:
        sy = pmb.DensityFilter(dom, radius=2.0)(sx)
        ss = pmb.SIMP(1e-9, 3)(sy)
        sK = pmb.AssembleStiffness(dom, bc=P.bc)(ss)
        mgs = pmb.solvers.auto_multigrid(dom, min_size=4)
        su = pmb.LinSolve(hermitian=True, solver=pmb.solvers.CG(preconditioner=mgs[0], tol=1e-8))(sK, P.f)
        sc = pmb.Compliance()(su, P.f)
    for it in range(2):
        c_ref = P.response(sx.state)
        dx_ref = P.sensitivity()
        assert isinstance(su.state, np.ndarray) and isinstance(sy.state, np.ndarray)
        assert abs(float(sc.state) - c_ref) <= 1e-6 * abs(c_ref)
        sc.sensitivity = 1.0
        fn.sensitivity()
        np.testing.assert_allclose(sx.sensitivity, dx_ref, rtol=1e-6, atol=1e-7 * np.abs(dx_ref).max())
        fn.reset()
        sx.state = np.clip(sx.state + 0.2 * (rng.random(dom.nel) - 0.5), 0, 1)
        fn.response()


def test_oc_update_and_minimize_oc_vs_reference(pmb):
    """BASELINE.json configs[0] driven by the OC update on the device: 10 iterations of the 2-D MBB 100x50 problem against
    the reference's own OC history (tests/golden/ref_oc_mbb100x50.npz: 369.3086 -> 101.4926, direct solver)."""
    import torch
    from pymoto_b200 import device as dv

    g = load("oc_mbb100x50")
    nx, ny = 100, 50
    dom = pmb.VoxelDomain(nx, ny)
    nodes = dom.nodes
    bc = np.concatenate([2 * nodes[0, :].flatten(), 2 * nodes[nx, 0].flatten() + 1])
    f = np.zeros(dom.nnodes * 2)
    f[2 * nodes[0, ny].flatten() + 1] = -1.0
    fd = dv.to_device(f)
    sx = pmb.Signal("x", state=dv.to_device(np.full(dom.nel, 0.5)))
    with pmb.Network() as fn:
        sy = pmb.DensityFilter(dom, radius=2.0)(sx)
        ss = pmb.SIMP(1e-9, 3)(sy)
        sK = pmb.AssembleStiffness(dom, bc=bc)(ss)
        mgs = pmb.solvers.auto_multigrid(dom)
        su = pmb.LinSolve(hermitian=True, solver=pmb.solvers.CG(preconditioner=mgs[0], tol=1e-10))(sK, fd)
        sc = pmb.Compliance()(su, fd)
    oc = pmb.OC(sx, sc, fn, verbosity=0)
    hist = []
    x = None
    for it in range(10):
        xnew, gval, dg = oc.step(x)
        hist.append(gval)
        x = xnew
        assert float(xnew.min()) >= 0.0 and float(xnew.max()) <= 1.0
        assert abs(float(xnew.mean()) - 0.5) < 1e-3  # volume kept by the bisection on the multiplier (l1l2tol = 1e-4)
    np.testing.assert_allclose(hist, g["history"], rtol=1e-6)
    np.testing.assert_allclose(x.cpu().numpy(), g["x10"], rtol=0, atol=1e-5)
    # one candidate evaluation against numpy
    xi, dgi = np.random.default_rng(0).random(1000), -np.random.default_rng(1).random(1000)
    out, xn = dv.empty(1), dv.empty(1000)
    from pymoto_b200 import _lib

    xid, dgid = dv.to_device(xi), dv.to_device(dgi)  # keep the tensors alive across the raw-pointer call
    _lib.call("pmb_oc_candidate", 1000, dv.ptr(xid), dv.ptr(dgid), 0.1, 0.0, 1.0, 0.37, dv.ptr(xn), dv.ptr(out),
              dv.ptr(dv.workspace().red), dv.stream())
    ref = np.clip(xi * np.sqrt(-dgi / 0.37), np.maximum(0.0, xi - 0.1), np.minimum(1.0, xi + 0.1))
    np.testing.assert_allclose(xn.cpu().numpy(), ref, rtol=1e-15)
    assert abs(float(out.item()) - ref.sum()) <= 1e-12 * ref.sum()
    # minimize_oc runs to its stopping criterion on a small problem
    oc2 = pmb.minimize_oc(sx, sc, function=fn, maxit=3, verbosity=0)
    assert oc2.iter <= 3


def test_linsolve_multiple_rhs_and_finite_difference(pmb):
    """reference tests/test_linsolve_sparse.py:32-91 style: K u = f for a block of right-hand sides (solved one by one
    through LDAS: the second, linearly dependent column costs no CG) and a finite-difference check of d(f.u)/dx."""
    import scipy.sparse.linalg as spla

    nx, ny, nz = 6, 4, 4
    dom = pmb.VoxelDomain(nx, ny, nz)
    P = ComplianceProblem(Grid(nx, ny, nz), kind="cantilever", tol=1e-10, min_size=2)
    rng = np.random.default_rng(4)
    x = rng.random(dom.nel) * 0.8 + 0.2
    asm = pmb.AssembleStiffness(dom, bc=P.bc)
    mgs = pmb.solvers.auto_multigrid(dom, min_size=2)
    cg = pmb.solvers.CG(preconditioner=mgs[0], tol=1e-11)
    ls = pmb.LinSolve(hermitian=True, solver=cg)
    F = np.stack([P.f, 2.5 * P.f, rng.standard_normal(P.f.size)], axis=1)
    F[P.bc, 2] = 0.0
    K = asm(x)
    U = ls(K, F)
    assert U.shape == F.shape
    Ks = K.tocsr().tocsc()
    for c in range(3):
        np.testing.assert_allclose(U[:, c], spla.spsolve(Ks, F[:, c]), rtol=0, atol=1e-8 * np.abs(U[:, c]).max())
    assert len(ls.solver.x_stored) == 2  # column 1 = 2.5 x column 0 came from the LDAS database
    # finite differences of c(x) = f . u(x) against the analytical sensitivity -lam_e^T Ke u_e (single rhs)
    ls1 = pmb.LinSolve(hermitian=True, solver=pmb.solvers.CG(preconditioner=pmb.solvers.auto_multigrid(dom, min_size=2)[0], tol=1e-12))
    u = ls1(asm(x), P.f)
    c0 = u @ P.f
    dmat, _ = ls1._sensitivity(P.f)
    dcdx = asm._sensitivity(dmat)[0]
    for e in rng.integers(0, dom.nel, 5):
        xp = x.copy()
        xp[e] += 1e-6
        lsp = pmb.LinSolve(hermitian=True, solver=pmb.solvers.CG(preconditioner=pmb.solvers.auto_multigrid(dom, min_size=2)[0], tol=1e-12))
        cp = lsp(asm(xp), P.f) @ P.f
        assert abs((cp - c0) / 1e-6 - dcdx[e]) <= 2e-4 * abs(dcdx[e]) + 1e-7


# ------------------------------------------------------------------------------------------------ MMA on the device (next row f3)
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["m1", "m2", "unconstrained", "m1_1987", "m3_vecbounds", "m5", "m6_vecbounds"])
def test_mma_device_update_vs_reference_golden(pmb, name):
    """pmb_mma_* kernels + the host Newton driver against pym.MMA.step on seeded subproblems (m = 1, 2, 3, 5, 6 constraints,
    unconstrained, MMA1987, vector bounds): asymptotes to rounding, new design within 2e-9 absolute."""
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden_opt_inputs import subsolv_inputs
    from pymoto_b200 import device as dv
    from pymoto_b200.optimizers import MmaDeviceOps, mma_design_update

    g = load("mma_subsolv")
    p = subsolv_inputs(name)
    n, m = p["n"], max(1, p["nresp"] - 1)
    ops = MmaDeviceOps(n, m)
    offset = dv.to_device(np.full(n, 0.5))
    bnd = {k: (dv.to_device(p[k]) if isinstance(p[k], np.ndarray) else p[k]) for k in ("xmin", "xmax", "move")}
    opt = dict(albefa=0.1, asyincr=1.2, asydecr=0.7, asybound=10.0, a0=1.0, epsimin=1e-10, rho=1e-5,
               version=1987 if "1987" in p["version"] else 2007, a=np.zeros(m), c=np.full(m, 1e3), d=np.ones(m))
    lam, its = mma_design_update(ops, dv.to_device(p["x"]), p["g"].copy(), [dv.to_device(r) for r in p["dg"]], offset,
                                 dv.to_device(p["xold1"]), dv.to_device(p["xold2"]), bnd["xmin"], bnd["xmax"], bnd["move"], opt)
    np.testing.assert_allclose(offset.cpu().numpy(), g[name + "_offset"], rtol=1e-15)
    np.testing.assert_allclose(ops.t["low"].cpu().numpy(), g[name + "_low"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(ops.t["upp"].cpu().numpy(), g[name + "_upp"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(ops.x.cpu().numpy(), g[name + "_xnew"], rtol=0, atol=2e-9)
    assert 5 < its < 200 and np.all(lam > 0)
    # the reductions are deterministic: a second solve of the same subproblem reproduces the design bit for bit
    first = ops.x.clone()
    offset2 = dv.to_device(np.full(n, 0.5))
    mma_design_update(ops, dv.to_device(p["x"]), p["g"].copy(), [dv.to_device(r) for r in p["dg"]], offset2,
                      dv.to_device(p["xold1"]), dv.to_device(p["xold2"]), bnd["xmin"], bnd["xmax"], bnd["move"], opt)
    assert bool((ops.x == first).all())


def _mma_chain(pmb, dom, bc, f, host):
    from pymoto_b200 import device as dv

    fd = f if host else dv.to_device(f)
    x0 = np.full(dom.nel, 0.5)
    sx = pmb.Signal("x", state=x0 if host else dv.to_device(x0))
    with pmb.Network() as fn:
        sy = pmb.DensityFilter(dom, radius=2.0)(sx)
        ss = pmb.SIMP(1e-9, 3)(sy)
        sK = pmb.AssembleStiffness(dom, bc=bc)(ss)
        mgs = pmb.solvers.auto_multigrid(dom)
        su = pmb.LinSolve(hermitian=True, solver=pmb.solvers.CG(preconditioner=mgs[0], tol=1e-10))(sK, fd)
        sc = pmb.Compliance()(su, fd)
        sg0 = pmb.Scaling(scaling=100.0)(sc)
        sv = pmb.Sum()(sy)
        sg1 = pmb.Scaling(scaling=10.0, maxval=0.5 * dom.nel)(sv)
    sg0.tag, sg1.tag = "objective", "volume constraint"
    return sx, [sg0, sg1], fn


@pytest.mark.gpu
@pytest.mark.parametrize("case,host", [("mma_mbb60x30", False), ("mma_hex16x8x8", False), ("mma_hex16x8x8", True)])
def test_mma_design_loop_vs_reference_history(pmb, case, host):
    """pymoto_b200.MMA driving the device chain (filter, SIMP, assembly, CG+GMG, compliance, scaled objective and volume
    constraint) against the reference's own MMA2007 history of the same problem (direct solver), design by design; both
    with CUDA-tensor Signals (resident loop) and with numpy Signals (reference glue compatible)."""
    g = load(case)
    if case == "mma_mbb60x30":
        nx, ny = 60, 30
        dom = pmb.VoxelDomain(nx, ny)
        nodes = dom.nodes
        bc = np.concatenate([2 * nodes[0, :].flatten(), 2 * nodes[nx, 0].flatten() + 1])
        f = np.zeros(dom.nnodes * 2)
        f[2 * nodes[0, ny].flatten() + 1] = -1.0
    else:
        nx, ny, nz = 16, 8, 8
        dom = pmb.VoxelDomain(nx, ny, nz)
        ndof, bc, f = cantilever(Grid(nx, ny, nz))
    sx, resp, fn = _mma_chain(pmb, dom, bc, f, host)
    mma = pmb.MMA(sx, resp, fn, verbosity=0)
    ghist, x, xs = [], None, []
    for it in range(len(g["ghist"])):
        xnew, gv, dg = mma.step(x)
        ghist.append(np.array(gv, dtype=float))
        x = xnew
        xs.append(x.cpu().numpy())
        assert mma.newton_iterations > 0
    np.testing.assert_allclose(np.array(ghist), g["ghist"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(xs[0], g["x1"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(xs[1], g["x2"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(xs[-1], g["xlast"], rtol=0, atol=1e-4)
    # asymptote offsets follow sign decisions on (x - xold1)(xold1 - xold2): allow a few flips where a variable barely moves
    assert np.mean(np.abs(mma.offset.cpu().numpy() / g["offset_last"] - 1.0) > 1e-9) < 0.01
    assert isinstance(sx.state, np.ndarray) == host
    # minimize_mma runs to a stopping criterion
    sx2, resp2, fn2 = _mma_chain(pmb, dom, bc, f, host)
    m2 = pmb.minimize_mma(sx2, resp2, function=fn2, maxit=2, verbosity=0)
    assert m2.iter <= 2


# ------------------------------------------------------------------------------------------------ VTI output of device fields (f4)
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["2d", "3d", "block"])
def test_write_to_vti_from_device_tensors(pmb, name, tmp_path):
    """Fields resident on the GPU are packed to Float32 by pmb_pack_f32 and written byte-for-byte like the reference's
    VoxelDomain.write_to_vti; mixing numpy and CUDA inputs is allowed; the WriteToVTI module numbers its files."""
    import sys

    import torch

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden_opt_inputs import vti_inputs

    shape, vecs, scale = vti_inputs(name)
    dom = pmb.VoxelDomain(*shape)
    want = load("vti")[name]
    dev = {k: (torch.from_numpy(np.ascontiguousarray(v)).cuda() if not np.iscomplexobj(v) else v) for k, v in vecs.items()}
    fn = str(tmp_path / "dev.vti")
    pmb.write_to_vti(dom, dev, fn, scale=scale)
    got = np.frombuffer(open(fn, "rb").read(), dtype=np.uint8)
    assert got.size == want.size and np.array_equal(got, want)
    if name == "3d":
        w = pmb.WriteToVTI(dom, str(tmp_path / "out" / "dat.vti"), scale=scale, interval=2)
        sigs = [pmb.Signal(k, state=v) for k, v in dev.items()]
        w(*sigs)
        w.response()  # iteration 1: skipped by the interval
        w.response()
        files = sorted(os.listdir(tmp_path / "out"))
        assert files == ["dat.0000.vti", "dat.0002.vti"], files
        assert np.array_equal(np.frombuffer(open(tmp_path / "out" / files[1], "rb").read(), dtype=np.uint8), want)


@pytest.mark.parametrize("n,m", [(200_000, 1), (60_001, 2)])
def test_mma_device_vs_oracle_larger_problem(pmb, n, m):
    """Three successive device MMA updates against the numpy oracle (oracle/nextrows.MMAOracle, itself pinned to the
    reference's fixtures) on a seeded problem larger than the fixtures: designs agree to 1e-7, asymptote offsets to rounding
    apart from sign flips of variables that did not move, the multi-block reductions (n > 592 x 256) are exercised."""
    from oracle.nextrows import MMAOracle
    from pymoto_b200 import device as dv
    from pymoto_b200.optimizers import MmaDeviceOps, mma_design_update

    rng = np.random.default_rng(n)
    wgt = 1.0 + rng.random(n)
    cons = [np.full(n, 1.0 / n)] + [rng.random(n) / n for _ in range(m - 1)]

    def responses(x):
        g = [np.sum(wgt / (x + 0.05)) / n] + [c @ x - 0.4 * c.sum() for c in cons]
        dg = [-wgt / (x + 0.05) ** 2 / n] + [c.copy() for c in cons]
        return np.array(g), np.array(dg)

    x0 = np.full(n, 0.4)
    o = MMAOracle(n, m + 1)
    ops = MmaDeviceOps(n, m)
    offset = dv.to_device(np.full(n, 0.5))
    opt = dict(albefa=0.1, asyincr=1.2, asydecr=0.7, asybound=10.0, a0=1.0, epsimin=1e-10, rho=1e-5, version=2007,
               a=np.zeros(m), c=np.full(m, 1e3), d=np.ones(m))
    xo, xd, xold1, xold2 = x0.copy(), dv.to_device(x0), None, None
    for it in range(3):
        g, dg = responses(xo)
        xo_new = o.step(xo.copy(), g, dg)
        g, dg = responses(xd.cpu().numpy())
        mma_design_update(ops, xd, g, [dv.to_device(r) for r in dg], offset, xold1, xold2, 0.0, 1.0, 0.1, opt)
        xold2, xold1 = xold1, xd.clone()
        xd, xo = ops.x.clone(), xo_new
        np.testing.assert_allclose(xd.cpu().numpy(), xo, rtol=0, atol=1e-7)
    assert np.mean(np.abs(offset.cpu().numpy() / o.offset - 1.0) > 1e-9) < 0.01


# ------------------------------------------------------------------------------------------------ C-side PCG driver (8b)
@pytest.mark.parametrize("shape,restart,matrix_free", [((16, 8, 8), 50, True), ((32, 16, 16), 3, True), ((32, 16, 16), 50, False)])
def test_c_pcg_driver_identical_to_python_driver(pmb, shape, restart, matrix_free):
    """pmb_pcg_solve / pmb_vcycle (whole solve driven from C, one host poll per iteration) issue the launches of the Python
    driver in the same order: solution bit-identical, same iteration count and residual, cold and warm start, explicit-
    residual and recurrence branches (restart = 3), matrix-free and CSR-streamed finest level."""
    import ctypes as C

    from pymoto_b200 import _lib, device as dv
    from pymoto_b200.matrix import DeviceCSR
    from pymoto_b200.solvers import CG

    nx, ny, nz = shape
    gr = Grid(nx, ny, nz)
    ndof, bc, f = cantilever(gr)
    dom = pmb.VoxelDomain(nx, ny, nz)
    rng = np.random.default_rng(11)
    s = 1e-9 + (1 - 1e-9) * rng.random(gr.nel) ** 3
    saved_mf, saved_c = DeviceCSR.matrix_free, CG.use_c_driver
    try:
        DeviceCSR.matrix_free = matrix_free
        K = pmb.AssembleStiffness(dom, bc=bc)(dv.to_device(s))
        mgs = pmb.solvers.auto_multigrid(dom, min_size=4)
        cg = CG(preconditioner=mgs[0], tol=1e-9, restart=restart)
        cg.update(K)
        fd = dv.to_device(f)
        x0 = dv.to_device(rng.standard_normal(f.size) * 1e-3)
        results = {}
        for use_c in (False, True):
            CG.use_c_driver = use_c
            if use_c:
                assert cg._mg_desc() is not None
            cold = cg.solve(fd).clone()
            its_cold, res_cold = cg.iterations, cg.last_residual
            warm = cg.solve(fd, x0=x0).clone()
            results[use_c] = (cold.cpu().numpy(), its_cold, res_cold, warm.cpu().numpy(), cg.iterations, cg.last_residual)
        py, c = results[False], results[True]
        assert py[1] == c[1] and py[4] == c[4], (py[1], c[1], py[4], c[4])
        # the C driver went through its plan: non-restart iterations were replayed as one captured graph
        assert cg._c_plan is not None
        if restart == 50 or py[1] + py[4] > 2 * (py[1] // restart + py[4] // restart + 2) + 2:
            assert _lib.query("pmb_pcg_plan_graph_replays", cg._c_plan) > 0
        assert py[1] > restart or restart == 50
        assert np.array_equal(py[0], c[0]) and np.array_equal(py[3], c[3])
        assert py[2] == c[2] and py[5] == c[5]
        Ks = K.tocsr()
        assert np.linalg.norm(Ks @ c[0] - f) <= 1e-8 * np.linalg.norm(f)
        # one V-cycle through pmb_vcycle against GeometricMultigrid.solve
        r = dv.to_device(rng.standard_normal(f.size))
        z_py = mgs[0].solve(r).clone()
        z_c = dv.empty(f.size)
        desc = cg._mg_desc()
        _lib.call("pmb_vcycle", C.byref(desc), dv.ptr(r), dv.ptr(z_c), dv.stream())
        assert np.array_equal(z_py.cpu().numpy(), z_c.cpu().numpy())
        # argument validation stays on the host
        desc.nlevels = 0
        with pytest.raises(_lib.PmbError):
            _lib.call("pmb_vcycle", C.byref(desc), dv.ptr(r), dv.ptr(z_c), dv.stream())
    finally:
        DeviceCSR.matrix_free, CG.use_c_driver = saved_mf, saved_c
