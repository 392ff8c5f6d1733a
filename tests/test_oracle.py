"""Pin the CPU oracle against the reference: committed golden fixtures (always) and the live reference (when
/root/reference exists).  Also the reference's own known-answer tests for this path, run on the oracle."""
import numpy as np
import pytest

import oracle
from oracle import Grid
from oracle.chain import ComplianceProblem
from _golden import CASES, DESIGNS, load, digest, force_vector
from _refimport import import_reference


@pytest.mark.parametrize("case", CASES)
def test_chain_against_golden(case):
    g = load(case)
    nx, ny, nz = (int(v) for v in g["shape"])
    P = ComplianceProblem(Grid(nx, ny, nz), kind=str(g["kind"]), radius=float(g["radius"]), xmin=float(g["xmin"]),
                          tol=float(g["tol"]), min_size=int(g["min_size"]))
    assert np.array_equal(P.bc, g["bc"])
    assert np.array_equal(P.f, force_vector(g, P.f.size))
    assert np.array_equal(P.asm.Ke, g["Ke"])  # bit-exact element matrix
    assert P.asm.bcdiagval == float(g["bcdiagval"])
    assert len(P.mgs) == int(g["n_mg"])
    for dname in DESIGNS:
        P.u = None  # every golden design was solved cold
        x = g[dname + "_x"]
        c = P.response(x)
        K = P.K
        assert K.nnz == int(g["nnz"])
        assert digest(K.indptr.astype(np.int32)) == str(g["indptr_sha256"])
        assert digest(K.indices.astype(np.int32)) == str(g["indices_sha256"])
        assert digest(K.data) == str(g[dname + "_data_sha256"])  # values bit-exact
        assert np.array_equal(P.y, g[dname + "_y"])  # filter bit-exact
        assert np.array_equal(K.diagonal(), g[dname + "_diag"])
        np.testing.assert_allclose(c, float(g[dname + "_compliance"]), rtol=1e-9)
        np.testing.assert_allclose(P.u, g[dname + "_u"], rtol=0, atol=1e-7 * np.abs(g[dname + "_u"]).max())
        dx = P.sensitivity()
        assert not P.solver.did_solve  # adjoint comes from the LDAS database, no CG
        np.testing.assert_allclose(dx, g[dname + "_dcdx"], rtol=1e-6, atol=1e-8 * np.abs(g[dname + "_dcdx"]).max())
        if dname == "random":
            assert np.array_equal(P.filt.sensitivity(g[dname + "_filter_bwd_in"]), g[dname + "_filter_bwd"])


def test_pattern_unique_equals_closed_form():
    for shape, ndof in [((5, 3, 2), 3), ((4, 3, 0), 2), ((3, 3, 3), 1), ((1, 1, 1), 3), ((2, 1, 0), 1)]:
        gr = Grid(*shape)
        ip1, ix1, dm1 = oracle.assembly.pattern_unique(gr, ndof)
        ip2, ix2 = oracle.assembly.pattern_closed_form(gr, ndof)
        assert np.array_equal(ip1, ip2) and np.array_equal(ix1, ix2)
        assert np.array_equal(dm1, oracle.assembly.datamap_closed_form(gr, ndof, ip2))


def test_transfer_against_golden():
    g = load("transfer")
    for name in ["2d", "3d", "3d1"]:
        shape = [int(v) for v in g[name + "_shape"]]
        ndof = int(g[name + "_ndof"])
        fine = Grid(*shape)
        R = oracle.solvers.prolongation_matrix(fine, fine.coarsen(), ndof)
        assert np.array_equal(R.T @ g[name + "_v"], g[name + "_restrict"])
        assert np.array_equal(R @ g[name + "_vc"], g[name + "_prolong"])
        assert np.array_equal(R.T @ np.ones(R.shape[0]), g[name + "_restrict_ones"])


def test_restriction_constants_reference_known_answers():
    """reference tests/test_solvers_multigrid.py:9-91: restriction of ones = 4/3/2.25 (2-D) and 8/6/4.5/3.375 (3-D),
    prolongation of ones = 1."""
    fine = Grid(8, 6)
    R = oracle.solvers.prolongation_matrix(fine, fine.coarsen(), 1)
    r = (R.T @ np.ones(fine.nnodes)).reshape(4, 5)  # (j, i)
    assert np.all(r[1:-1, 1:-1] == 4.0) and r[0, 0] == 2.25 and np.all(r[0, 1:-1] == 3.0) and np.all(r[1:-1, 0] == 3.0)
    assert np.all(R @ np.ones(R.shape[1]) == 1.0)
    fine = Grid(4, 6, 8)
    R = oracle.solvers.prolongation_matrix(fine, fine.coarsen(), 1)
    r = (R.T @ np.ones(fine.nnodes)).reshape(5, 4, 3)  # (k, j, i)
    assert np.all(r[1:-1, 1:-1, 1:-1] == 8.0) and r[0, 0, 0] == 3.375
    assert np.all(r[0, 1:-1, 1:-1] == 6.0) and np.all(r[0, 0, 1:-1] == 4.5)
    assert np.all(R @ np.ones(R.shape[1]) == 1.0)


def test_single_element_equals_Ke():
    """reference tests/test_assembly.py:22-35."""
    gr = Grid(1, 1, 1)
    Ke = oracle.assembly.stiffness_element(gr)
    K = oracle.assembly.Assembler(gr, Ke)(np.array([1.0]))
    assert np.array_equal(K.toarray(), Ke)


def test_bc_rows_cols(): 
    """reference tests/test_assembly.py:37-61: bc rows/cols zero, diagonal = bcdiagval."""
    gr = Grid(3, 2, 2)
    Ke = oracle.assembly.stiffness_element(gr)
    bc = np.array([0, 4, 17, 50])
    K = oracle.assembly.Assembler(gr, Ke, bc=bc, bcdiagval=7.5)(np.random.default_rng(0).random(gr.nel)).toarray()
    for b in bc:
        row = K[b].copy(); col = K[:, b].copy()
        assert row[b] == 7.5
        row[b] = 0; col[b] = 0
        assert not row.any() and not col.any()


def test_jacobi_cg_against_golden():
    g = load("jacobi_cg")
    gr = Grid(*[int(v) for v in g["shape"]])
    K = oracle.assembly.Assembler(gr, oracle.assembly.stiffness_element(gr), bc=g["bc"])(g["x"])
    jac = oracle.solvers.DampedJacobi(w=1.0)
    cg = oracle.solvers.CG(jac, tol=1e-10)
    cg.update(K)
    u = cg.solve(g["f"])
    np.testing.assert_allclose(u, g["u"], rtol=0, atol=1e-9 * np.abs(g["u"]).max())


def test_live_reference_random_grids():
    """Oracle vs the live reference on shapes not in the fixtures (skipped where /root/reference is absent)."""
    pym = import_reference()
    if pym is None:
        pytest.skip("reference not available")
    rng = np.random.default_rng(11)
    for shape, ndof in [((5, 4, 3), 3), ((7, 5, 0), 2), ((4, 4, 5), 1)]:
        gr = Grid(*shape)
        d = pym.VoxelDomain(*shape)
        x = rng.random(gr.nel)
        bc = np.unique(rng.integers(0, gr.nnodes * ndof, 9))
        if ndof == 1:
            mod = pym.AssemblePoisson(d, bc=bc)
            Ke = oracle.assembly.poisson_element(gr)
        else:
            mod = pym.AssembleStiffness(d, bc=bc)
            Ke = oracle.assembly.stiffness_element(gr)
        assert np.array_equal(Ke, mod.elmat[0])
        Kr = mod(x)
        Ko = oracle.assembly.Assembler(gr, Ke, bc=bc)(x)
        assert np.array_equal(Kr.indptr, Ko.indptr) and np.array_equal(Kr.indices, Ko.indices)
        assert np.array_equal(Kr.data, Ko.data)
        fr = pym.DensityFilter(d, radius=2.5)
        fo = oracle.filter.DensityFilter(gr, radius=2.5)
        assert np.array_equal(fr(x), fo(x))
        assert np.array_equal(fr._sensitivity(x), fo.sensitivity(x))


def test_filterconv_oracle_against_golden():
    """Oracle FilterConv (scipy.signal, padded index array) vs the reference's outputs for every boundary mode."""
    from oracle.nextrows import FilterConv

    g = load("filterconv")
    cases = {
        "sym3d": ((7, 5, 4), dict(radius=2.0)),
        "r3_3d": ((9, 6, 5), dict(radius=3.2)),
        "sym2d": ((12, 9, 0), dict(radius=2.5)),
        "edge_wrap": ((8, 6, 5), dict(radius=2.0, xmin_bc="edge", xmax_bc="wrap", ymin_bc="wrap", ymax_bc="wrap", zmin_bc="edge", zmax_bc="edge")),
        "const": ((6, 7, 5), dict(radius=2.0, xmin_bc=0.0, xmax_bc=1.0, ymin_bc=0.25, ymax_bc="symmetric", zmin_bc="edge", zmax_bc=0.75)),
        "weights": ((6, 5, 4), dict(xmin_bc="wrap", xmax_bc="symmetric", ymin_bc=0.5, ymax_bc="edge")),
        "weights2d": ((9, 8, 0), dict(xmin_bc="edge", ymax_bc=2.0)),
        "override": ((6, 6, 4), dict(radius=2.0)),
    }
    for name, (shape, kw) in cases.items():
        if name + "_weights" in g.files:
            kw = dict(kw, weights=g[name + "_weights"])
        f = FilterConv(Grid(*shape), **kw)
        if name == "override":
            f.override_values((np.s_[1:3], np.s_[2:4], np.s_[:]), 1.0)
        x = g[name + "_x"]
        np.testing.assert_allclose(f(x), g[name + "_y"], rtol=0, atol=1e-13 * np.abs(g[name + "_y"]).max())
        np.testing.assert_allclose(f.sensitivity(g[name + "_dy"], x.size), g[name + "_dx"], rtol=0, atol=1e-13 * np.abs(g[name + "_dx"]).max())


def test_oc_oracle_against_golden():
    """Oracle OC update + chain (direct solver) reproduces the reference's 10-iteration history of the 2-D MBB problem."""
    import scipy.sparse.linalg as spla
    from oracle.nextrows import oc_update

    g = load("oc_mbb100x50")
    nx, ny = 100, 50
    gr = Grid(nx, ny)
    nodes = gr.nodes3d()
    bc = np.concatenate([2 * nodes[0, :].ravel(), 2 * nodes[nx, 0].ravel() + 1])
    f = np.zeros(gr.nnodes * 2)
    f[2 * nodes[0, ny].ravel() + 1] = -1.0
    flt = oracle.filter.DensityFilter(gr, 2.0)
    asm = oracle.assembly.Assembler(gr, oracle.assembly.stiffness_element(gr), bc=bc)
    x = np.full(gr.nel, 0.5)
    hist = []
    for it in range(4):  # four iterations are enough to pin the update rule
        y = flt(x)
        K = asm(1e-9 + (1 - 1e-9) * y ** 3)
        u = spla.spsolve(K.tocsc(), f)
        hist.append(u @ f)
        ds = asm.sensitivity(-u, u)
        dx = flt.sensitivity(ds * 3 * (1 - 1e-9) * y ** 2)
        x = oc_update(x, dx)
    np.testing.assert_allclose(hist, g["history"][:4], rtol=1e-9)


@pytest.mark.parametrize("name", ["m1", "m2", "unconstrained", "m1_1987", "m3_vecbounds", "m5", "m6_vecbounds"])
def test_mma_oracle_against_golden(name):
    """Oracle MMA update (numpy restatement of pymoto/common/mma.py) against pym.MMA.step on seeded subproblems."""
    from make_golden_opt_inputs import subsolv_inputs
    from oracle.nextrows import MMAOracle

    g = load("mma_subsolv")
    p = subsolv_inputs(name)
    o = MMAOracle(p["n"], p["nresp"], move=p["move"], xmin=p["xmin"], xmax=p["xmax"], version=p["version"])
    o.xold1, o.xold2 = p["xold1"].copy(), p["xold2"].copy()
    xnew = o.step(p["x"].copy(), p["g"], p["dg"])
    np.testing.assert_allclose(o.offset, g[name + "_offset"], rtol=1e-15)
    np.testing.assert_allclose(o.low, g[name + "_low"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(xnew, g[name + "_xnew"], rtol=0, atol=1e-10)


@pytest.mark.parametrize("name", ["hex_two_const", "quad_const", "quad_three_bc", "hex_two_thermal"])
def test_assembly_several_matrices_and_constant_against_golden(name):
    """Oracle Assembler with a list of element matrices / add_constant against pym.AssembleGeneral (assembly.py:245-253, 294-295):
    the same bits for the matrix (the scaled element matrices are summed before the scatter, like the reference does), the
    dyad contraction per element matrix to rounding."""
    from make_golden_opt_inputs import asm_inputs

    g = load("asm_multi")
    p = asm_inputs(name)
    gr = Grid(*p["shape"])
    asm = oracle.assembly.Assembler(gr, p["mats"], bc=p["bc"], add_constant=p["const"])
    K = asm(*p["xs"])
    assert np.array_equal(K.toarray(), g[name + "_K"])
    dx = asm.sensitivity(p["u"], p["v"])
    for i, d in enumerate(dx if isinstance(dx, list) else [dx]):
        np.testing.assert_allclose(d, g[f"{name}_dx{i}"], rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("name", ["gcmma_m2", "gcmma_unconstrained"])
def test_gcmma_oracle_against_golden(name):
    """The numpy GCMMA restatement (inner iterations, rho initialisation / update, conservative-approximation test) against six
    outer iterations of pym.MMA(mmaversion="GCMMA") on the seeded non-convex problems: same number of response evaluations."""
    from make_golden_opt_inputs import GCMMA_CASES, gcmma_problem
    from oracle.nextrows import MMAOracle

    g = load("gcmma")
    n, x0, responses = gcmma_problem(name)
    opt = MMAOracle(n, GCMMA_CASES[name][1], version="GCMMA")
    x = x0.copy()
    for it in range(6):
        count = [0]

        def evaluate(xc):
            count[0] += 1
            return responses(xc)[0]

        gk, dg = responses(x)
        x = opt.step(x, gk, dg, evaluate=evaluate)
        assert count[0] == int(g[name + "_nev"][it])
        np.testing.assert_allclose(opt.rho, g[name + "_rho"][it], rtol=1e-8)
        np.testing.assert_allclose(opt.g_last, g[name + "_g"][it], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(x, g[name + "_x"][it], rtol=0, atol=1e-7)
    np.testing.assert_allclose(opt.offset, g[name + "_offset"], rtol=1e-13)
