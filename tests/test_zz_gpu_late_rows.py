"""GPU tests of the rows finished last (GCMMA, slice_network, several element matrices / add_constant in the assembly).
Kept in a file that sorts after the parity suite so that `pytest -x` reaches these only once every hot-path parity test passed."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))


@pytest.fixture(scope="module")
def pmb():
    import torch
    import pymoto_b200 as pmb

    assert torch.cuda.is_available()
    return pmb


# ------------------------------------------------------------------------------------------------ GCMMA (mma.py:104-160)
@pytest.mark.parametrize("name", ["gcmma_m2", "gcmma_unconstrained"])
def test_gcmma_device_passes_vs_reference_history(pmb, name):
    """pmb_mma_gcmma_rho / pmb_mma_gcmma_estimate + the inner-iteration driver against six outer iterations of the reference's
    GCMMA on a non-convex analytic problem: same number of response evaluations in every outer iteration, same rho, g, designs."""
    from make_golden_opt_inputs import GCMMA_CASES
    from pymoto_b200 import device as dv
    from pymoto_b200.optimizers import MmaDeviceOps
    from test_mma_cpu import check_gcmma_history, gcmma_history

    n, nresp = GCMMA_CASES[name]
    ops = MmaDeviceOps(n, max(1, nresp - 1))
    ops.from_numpy = lambda a: dv.to_device(np.ascontiguousarray(a, dtype=float))
    check_gcmma_history(name, gcmma_history(ops, name, 6))


@pytest.mark.parametrize("host", [True, False])
def test_gcmma_class_through_network_vs_reference_history(pmb, host):
    """pymoto_b200.MMA(mmaversion="GCMMA") driving a Network (responses re-evaluated at every inner candidate, sensitivities
    once per outer iteration) against the same reference history; slice_network leaves the unrelated module alone."""
    import torch
    from _golden import load
    from make_golden_opt_inputs import gcmma_problem

    name = "gcmma_m2"
    n, x0, responses = gcmma_problem(name)
    g = load("gcmma")
    evals, other = [0], [0]

    def to_np(v):
        return v.cpu().numpy() if torch.is_tensor(v) else np.asarray(v)

    class Analytic(pmb.Module):
        def __call__(self, x):
            evals[0] += 1
            return tuple(float(v) for v in responses(to_np(x))[0])

        def _sensitivity(self, *dg):
            J = responses(to_np(self.sig_in[0].state))[1]
            d = sum(float(di) * J[i] for i, di in enumerate(dg) if di is not None)
            return d if host else torch.as_tensor(d, device="cuda")

    class Unrelated(pmb.Module):
        def __call__(self, x):
            other[0] += 1
            return to_np(x).sum()

    sx = pmb.Signal("x", state=x0.copy() if host else torch.as_tensor(x0, device="cuda"))
    fn = pmb.Network()
    with fn:
        resp = list(Analytic()(sx))
        Unrelated()(sx)
    mma = pmb.MMA(sx, resp, fn, verbosity=0, mmaversion="GCMMA", slice_network=True)
    x = mma.x
    other[0] = 0
    for it in range(6):
        e0 = evals[0]
        xnew, gv, dg = mma.step(x)
        assert evals[0] - e0 == int(g[name + "_nev"][it]) and mma.gcmma_inner_iterations >= 1
        np.testing.assert_allclose(gv, g[name + "_g"][it], rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(mma.rho, g[name + "_rho"][it], rtol=1e-6)
        np.testing.assert_allclose(xnew.cpu().numpy(), g[name + "_x"][it], rtol=0, atol=2e-6)
        x = xnew
    assert other[0] == 0  # slice_network: the module the responses do not depend on never ran
    assert isinstance(sx.state, np.ndarray) == host


# ------------------------------------------------------------------------------------------------ several element matrices / add_constant
@pytest.mark.parametrize("name", ["hex_two_const", "quad_const", "quad_three_bc", "hex_two_thermal"])
def test_assembly_several_matrices_and_constant_vs_reference(pmb, name):
    """pymoto_b200.AssembleGeneral with a list of element matrices, Dirichlet dofs and add_constant against the unmodified
    reference's output on seeded inputs (assembly.py:245-253, 294-295).  The device sums separately assembled matrices, the
    reference sums scaled element matrices before scattering: equal to rounding (1e-14 of the largest entry), not bit for bit."""
    import torch
    from _golden import load
    from make_golden_opt_inputs import asm_inputs

    g = load("asm_multi")
    p = asm_inputs(name)
    dom = pmb.VoxelDomain(*p["shape"])
    mats = p["mats"] if len(p["mats"]) > 1 else p["mats"][0]
    for device_inputs in (False, True):
        asm = pmb.AssembleGeneral(dom, mats, bc=p["bc"], add_constant=p["const"])
        sigs = [pmb.Signal(f"x{i}", state=(torch.as_tensor(x, device="cuda") if device_inputs else x.copy())) for i, x in enumerate(p["xs"])]
        sK = asm(*sigs)
        K = sK.state
        assert K.generator is None  # applied from the assembled values on every level
        want = g[name + "_K"]
        np.testing.assert_allclose(K.toarray(), want, rtol=0, atol=1e-14 * np.abs(want).max())
        # the operator products use the summed values (row statistics re-derived from them)
        xv = np.random.default_rng(1).standard_normal(K.shape[0])
        np.testing.assert_allclose((K @ xv), want @ xv, rtol=0, atol=1e-12 * np.abs(want @ xv).max())
        np.testing.assert_allclose(K.diagonal(), np.diag(want), rtol=0, atol=1e-14 * np.abs(want).max())
        sK.sensitivity = pmb.DeviceDyad(torch.as_tensor(p["u"], device="cuda"), torch.as_tensor(p["v"], device="cuda"))
        asm.sensitivity()
        for i, s in enumerate(sigs):
            assert torch.is_tensor(s.sensitivity) == device_inputs
            got = s.sensitivity.cpu().numpy() if device_inputs else s.sensitivity
            np.testing.assert_allclose(got, g[f"{name}_dx{i}"], rtol=1e-12, atol=1e-13)
    with pytest.raises(ValueError):
        asm(*sigs, sigs[0])
    if len(p["mats"]) > 1:
        with pytest.raises(ValueError):
            pmb.AssembleGeneral(dom, [p["mats"][0], p["mats"][1][:-1]])


def test_two_material_solve_with_constant_vs_scipy(pmb):
    """A two-material stiffness (two element matrices, two density fields) plus a constant diagonal spring matrix solved with
    CG + geometric multigrid on assembled values (no matrix-free level 0, two-pass Galerkin) against scipy on the oracle's matrix;
    sensitivities of the compliance to both fields against the oracle's adjoint expressions."""
    import scipy.sparse as sps
    import scipy.sparse.linalg as spla

    import oracle
    from oracle import Grid
    from oracle.chain import cantilever

    gr = Grid(16, 8, 8)
    ndof, bc, f = cantilever(gr)
    dom = pmb.VoxelDomain(16, 8, 8)
    rng = np.random.default_rng(5)
    K1, K2 = oracle.assembly.stiffness_element(gr), oracle.assembly.stiffness_element(gr, e_modulus=0.3, poisson_ratio=0.2)
    x1, x2 = 0.2 + 0.8 * rng.random(gr.nel), 0.1 + rng.random(gr.nel)
    spring = sps.diags(1e-3 * rng.random(gr.nnodes * 3)).tocsr()
    Ko = oracle.assembly.Assembler(gr, [K1, K2], bc=bc, add_constant=spring)
    Kref = Ko(x1, x2)
    uref = spla.spsolve(Kref.tocsc(), f)
    s1, s2 = pmb.Signal("x1", state=x1.copy()), pmb.Signal("x2", state=x2.copy())
    with pmb.Network() as fn:
        asm = pmb.AssembleGeneral(dom, [K1, K2], bc=bc, add_constant=spring)
        sK = asm(s1, s2)
        cg = pmb.solvers.CG(preconditioner=pmb.solvers.auto_multigrid(dom, min_size=2)[0], tol=1e-10)
        su = pmb.LinSolve(hermitian=True, solver=cg)(sK, f)
        sc = pmb.Compliance()(su, f)
    u = su.state
    assert np.linalg.norm(Kref @ u - f) <= 1e-8 * np.linalg.norm(f)
    np.testing.assert_allclose(u, uref, rtol=0, atol=1e-7 * np.abs(uref).max())
    assert 0 < cg.iterations < 40
    sc.sensitivity = 1.0
    fn.sensitivity()
    d1, d2 = Ko.sensitivity(-uref, uref)  # dc/dx_i = -u_e^T K_i u_e
    np.testing.assert_allclose(s1.sensitivity, d1, rtol=1e-6, atol=1e-8 * np.abs(d1).max())
    np.testing.assert_allclose(s2.sensitivity, d2, rtol=1e-6, atol=1e-8 * np.abs(d2).max())
