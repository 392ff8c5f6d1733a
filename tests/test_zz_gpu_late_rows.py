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
