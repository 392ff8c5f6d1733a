"""CPU tests of the MMA row and the VTI writer (8f row 4) -- no GPU needed.

The product's Newton driver (pymoto_b200.optimizers.mma_subsolv / mma_design_update) is backend-agnostic: on the GPU its
n-sized passes are the pmb_mma_* kernels; here they are host loops over the SAME per-variable arithmetic header
(pymoto_b200/csrc/pmb_mma_math.h) built by tests/mma_host_harness.cpp.  That pins the formulas and the driver against
the reference's own MMA (golden fixtures from tests/golden/make_golden_opt.py) before the kernels run on a device; the
GPU tests then check the kernels against the same fixtures.  The harness is test infrastructure and is never used by
pymoto_b200 itself.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

from _golden import load  # noqa: E402

NAMES = ("x", "xsi", "eta", "xo", "xsio", "etao", "dx", "dxsi", "deta", "low", "upp", "alfa", "beta", "P", "Q")


class HVecs(C.Structure):
    _fields_ = [(nm, C.c_void_p) for nm in NAMES]


class HBound(C.Structure):
    _fields_ = [("s", C.c_double), ("v", C.c_void_p)]


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = tmp_path_factory.mktemp("mma") / "libmma_host.so"
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", os.path.join(HERE, "mma_host_harness.cpp"), "-o", str(out)],
                   check=True)
    return C.CDLL(str(out))


class HostOps:
    """Same interface as pymoto_b200.optimizers.MmaDeviceOps on numpy arrays (test infrastructure)."""

    def __init__(self, lib, n, m):
        self.lib, self.n, self.m = lib, n, m
        self.t = {nm: np.zeros(n * (m + 1) if nm in ("P", "Q") else n) for nm in NAMES}
        self.vecs = HVecs(*[self.t[nm].ctypes.data for nm in NAMES])
        self.out = np.zeros(64)

    @property
    def x(self):
        return self.t["x"]

    @staticmethod
    def zeros(n):
        return np.zeros(n)

    @staticmethod
    def _h(vals):
        return (C.c_double * len(vals))(*[float(v) for v in vals])

    @staticmethod
    def _b(b):
        return HBound(0.0, b.ctypes.data) if isinstance(b, np.ndarray) else HBound(float(b), None)

    @staticmethod
    def _p(a):
        return C.c_void_p(a.ctypes.data)

    def asymptotes(self, x, xold1, xold2, offset, incr, decr, bound):
        assert self.lib.hmma_asymptotes(C.c_longlong(self.n), self._p(x), self._p(xold1), self._p(xold2), C.c_double(incr), C.c_double(decr),
                                        C.c_double(bound), self._p(offset)) == 0

    def setup(self, xval, dg_rows, offset, xmin, xmax, move, albefa, rho, version):
        rows = (C.c_void_p * (self.m + 1))(*[r.ctypes.data for r in dg_rows])
        assert self.lib.hmma_setup(C.c_longlong(self.n), self.m, self._p(xval), rows, self._p(offset), self._b(xmin), self._b(xmax),
                                   self._b(move), C.c_double(albefa), self._h(rho), int(version), C.byref(self.vecs), self._p(self.out)) == 0
        return self.out[: self.m + 1].copy()

    def _res(self):
        o = self.out
        return float(o[0]), o[1: self.m + 1].copy(), float(o[self.m + 1])

    def residual(self, lam, epsi):
        assert self.lib.hmma_residual(C.c_longlong(self.n), self.m, C.byref(self.vecs), self._h(lam), C.c_double(epsi), self._p(self.out)) == 0
        return self._res()

    def newton_sums(self, lam, epsi):
        m = self.m
        assert self.lib.hmma_newton_sums(C.c_longlong(self.n), m, C.byref(self.vecs), self._h(lam), C.c_double(epsi), self._p(self.out)) == 0
        o = self.out
        return o[:m].copy(), o[m: 2 * m].copy(), o[2 * m: 2 * m + m * m].reshape(m, m).copy()

    def newton_dir(self, lam, dlam, epsi):
        assert self.lib.hmma_newton_dir(C.c_longlong(self.n), self.m, C.byref(self.vecs), self._h(lam), self._h(dlam), C.c_double(epsi),
                                        self._p(self.out)) == 0
        return self.out[1:5].copy()

    def rho_sums(self, dg_rows, xmin, xmax):
        rows = (C.c_void_p * (self.m + 1))(*[r.ctypes.data for r in dg_rows])
        assert self.lib.hmma_gcmma_rho(C.c_longlong(self.n), self.m, rows, self._b(xmin), self._b(xmax), self._p(self.out)) == 0
        return self.out[: self.m + 1].copy()

    def estimate(self, xval, xmin, xmax):
        assert self.lib.hmma_gcmma_estimate(C.c_longlong(self.n), self.m, C.byref(self.vecs), self._p(xval), self._b(xmin), self._b(xmax),
                                            self._p(self.out)) == 0
        return self.out[: self.m + 1].copy(), float(self.out[self.m + 1])

    def linesearch(self, lam, steg, epsi):
        assert self.lib.hmma_linesearch(C.c_longlong(self.n), self.m, C.byref(self.vecs), self._h(lam), C.c_double(steg), C.c_double(epsi),
                                        self._p(self.out)) == 0
        return self._res()


DEFAULTS = dict(albefa=0.1, asyincr=1.2, asydecr=0.7, asybound=10.0, a0=1.0, epsimin=1e-10, rho=1e-5)


def run_update(ops_factory, p):
    from pymoto_b200.optimizers import mma_design_update

    n, nresp = p["n"], p["nresp"]
    m = max(1, nresp - 1)
    ops = ops_factory(n, m)
    offset = np.full(n, 0.5)
    opt = dict(DEFAULTS, version=1987 if "1987" in p["version"] else 2007, a=np.zeros(m), c=np.full(m, 1e3), d=np.ones(m))
    lam, its = mma_design_update(ops, p["x"].copy(), p["g"].copy(), [r.copy() for r in p["dg"]], offset, p["xold1"], p["xold2"],
                                 p["xmin"], p["xmax"], p["move"], opt)
    return ops, offset, lam, its


@pytest.mark.parametrize("name", ["m1", "m2", "unconstrained", "m1_1987", "m3_vecbounds", "m5", "m6_vecbounds"])
def test_mma_update_host_arithmetic_vs_reference_golden(harness, name):
    """asymptotes + set-up + primal-dual Newton solve against pym.MMA.step on the same seeded subproblem."""
    from make_golden_opt_inputs import subsolv_inputs

    g = load("mma_subsolv")
    p = subsolv_inputs(name)
    ops, offset, lam, its = run_update(lambda n, m: HostOps(harness, n, m), p)
    np.testing.assert_allclose(offset, g[name + "_offset"], rtol=1e-15)
    np.testing.assert_allclose(ops.t["low"], g[name + "_low"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(ops.t["upp"], g[name + "_upp"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(ops.x, g[name + "_xnew"], rtol=0, atol=2e-9)
    assert 5 < its < 200 and np.all(lam > 0)


def test_mma_live_reference_multi_iteration(harness):
    """Where the reference is importable: five successive updates on an analytic problem, design by design."""
    from _refimport import import_reference
    from pymoto_b200.optimizers import mma_design_update

    pym = import_reference()
    if pym is None:
        pytest.skip("reference not available on this machine (fixtures cover it)")
    n = 120
    rng = np.random.default_rng(3)
    w = 1.0 + rng.random(n)

    def resp(x):  # weighted compliance-like objective and a volume constraint
        return np.array([np.sum(w / (x + 0.05)), np.sum(x) / n - 0.4]), np.vstack([-w / (x + 0.05) ** 2, np.full(n, 1.0 / n)])

    x0 = np.full(n, 0.4)
    sig = pym.Signal("x", state=x0.copy())
    ref = pym.MMA([sig], [pym.Signal("g0", state=1.0), pym.Signal("g1", state=0.0)], pym.Network(), verbosity=0)
    ops = HostOps(harness, n, 1)
    offset = np.full(n, 0.5)
    opt = dict(DEFAULTS, version=2007, a=np.zeros(1), c=np.full(1, 1e3), d=np.ones(1))
    x, xr, xold1, xold2 = x0.copy(), x0.copy(), None, None
    for _ in range(5):
        g, dg = resp(xr)
        xr_new, _, _ = ref.step(xr.copy(), g.copy(), dg.copy())
        g, dg = resp(x)
        mma_design_update(ops, x, g, [dg[0].copy(), dg[1].copy()], offset, xold1, xold2, 0.0, 1.0, 0.1, opt)
        xold2, xold1 = xold1, x.copy()
        x, xr = ops.x.copy(), xr_new.copy()
        np.testing.assert_allclose(x, xr, rtol=0, atol=5e-8)
        np.testing.assert_allclose(offset, ref.offset, rtol=1e-14)


def gcmma_history(ops, name, iters):
    """Outer GCMMA iterations of a seeded analytic problem with the product's driver; shared with the GPU test."""
    from make_golden_opt_inputs import gcmma_problem
    from pymoto_b200.optimizers import gcmma_design_update

    n, x0, responses = gcmma_problem(name)
    as_np = (lambda a: a) if isinstance(ops.x, np.ndarray) else (lambda a: a.cpu().numpy())
    to_ops = (lambda a: np.ascontiguousarray(a)) if isinstance(ops.x, np.ndarray) else ops.from_numpy
    offset = to_ops(np.full(n, 0.5))
    opt = dict(DEFAULTS, version=2007, a=np.zeros(ops.m), c=np.full(ops.m, 1e3), d=np.ones(ops.m))
    x, xold1, xold2, xs, gs, rhos, nev = x0.copy(), None, None, [], [], [], []
    for _ in range(iters):
        gk, dg = responses(x)
        count = [0]

        def evaluate(xc):
            count[0] += 1
            return responses(as_np(xc))[0]

        xd = to_ops(x)
        if xold1 is not None and xold2 is not None:
            ops.asymptotes(xd, to_ops(xold1), to_ops(xold2), offset, opt["asyincr"], opt["asydecr"], opt["asybound"])
        r = gcmma_design_update(ops, xd, gk, [to_ops(row) for row in dg], offset, 0.0, 1.0, 0.1, opt, evaluate, maxit=20)
        xold2, xold1 = xold1, x.copy()
        x = as_np(ops.x).copy()
        xs.append(x); gs.append(r["g"]); rhos.append(r["rho"]); nev.append(count[0])
    return np.array(xs), np.array(gs), np.array(rhos), np.array(nev), as_np(offset)


def check_gcmma_history(name, got):
    g = load("gcmma")
    xs, gs, rhos, nev, offset = got
    assert np.array_equal(nev, g[name + "_nev"])  # same number of inner iterations in every outer iteration
    np.testing.assert_allclose(rhos, g[name + "_rho"], rtol=1e-6)
    np.testing.assert_allclose(gs, g[name + "_g"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(xs[0], g[name + "_x"][0], rtol=0, atol=1e-7)
    np.testing.assert_allclose(xs, g[name + "_x"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(offset, g[name + "_offset"], rtol=1e-12)


@pytest.mark.parametrize("name", ["gcmma_m2", "gcmma_unconstrained"])
def test_gcmma_host_arithmetic_vs_reference_history(harness, name):
    """GCMMA: the product's inner-iteration driver (rho initialisation / update, conservative-approximation test) on the host
    harness against six outer iterations of pym.MMA(mmaversion="GCMMA") on a non-convex analytic problem."""
    from make_golden_opt_inputs import GCMMA_CASES

    n, nresp = GCMMA_CASES[name]
    check_gcmma_history(name, gcmma_history(HostOps(harness, n, max(1, nresp - 1)), name, 6))


def test_network_cones_and_slice_selection():
    """slice_network (pymoto/common/optimizers.py:55-72): only the modules between the variables and the responses run."""
    import pymoto_b200 as pmb
    from pymoto_b200.optimizers import select_network

    calls = []

    def mod(tag, fn):
        class M(pmb.Module):
            def __call__(self, *a):
                calls.append(tag)
                return fn(*a)

        M.__name__ = tag
        return M()

    sx, sother = pmb.Signal("x", state=np.arange(4.0)), pmb.Signal("o", state=np.ones(3))
    fn = pmb.Network()
    with fn:
        sa = mod("A", lambda x: 2 * x)(sx)
        sb = mod("B", lambda o: o + 1)(sother)          # does not depend on x
        sg = mod("G", lambda a: float(a.sum()))(sa)
        sunused = mod("U", lambda a: a * 3)(sa)          # depends on x, response does not depend on it
        sh = mod("H", lambda a, b: float(a.sum() + b.sum()))(sa, sb)
    assert [type(m).__name__ for m in fn.get_input_cone(sx)] == ["A", "G", "U", "H"]
    assert [type(m).__name__ for m in fn.get_output_cone(sg)] == ["A", "G"]
    sub = select_network(fn, [sx], [sg], slice_network=True)
    assert [type(m).__name__ for m in sub] == ["A", "G"]
    assert [type(m).__name__ for m in select_network(fn, [sx], [sh], True)] == ["A", "H"]  # B feeds H but does not depend on x
    assert select_network(fn, [sx], [sg], False) is fn
    calls.clear()
    sub.response()
    assert calls == ["A", "G"]
    with pytest.raises(RuntimeError):
        select_network(fn, [sother], [sg], True)
    with fn:
        assert select_network(None, [sx], [sg]) is fn
    if not pmb.core.HAVE_PYMOTO:
        with pytest.raises(RuntimeError):
            select_network(None, [sx], [sg])


def test_mma_constructor_contract_without_gpu():
    """No CPU fallback: the product MMA refuses to construct without CUDA; option validation mirrors the reference."""
    import torch

    import pymoto_b200 as pmb

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(pmb.PmbError):
        pmb.MMA(pmb.Signal("x", state=np.ones(4)), [pmb.Signal("g", state=1.0)], None)
    with pytest.raises(pmb.PmbError):
        pmb.optimizers.MmaDeviceOps(10, 1)


# ------------------------------------------------------------------------------------------------ VTI (host arrays)
@pytest.mark.parametrize("name", ["2d", "3d", "block"])
def test_write_to_vti_bytes_equal_reference(tmp_path, name):
    """pymoto_b200.write_to_vti on numpy inputs writes byte-for-byte what VoxelDomain.write_to_vti of the reference writes."""
    from make_golden_opt_inputs import vti_inputs
    from pymoto_b200.domain import VoxelDomain
    from pymoto_b200.io import write_to_vti

    shape, vecs, scale = vti_inputs(name)
    fn = str(tmp_path / "out")  # extension appended like the reference does
    write_to_vti(VoxelDomain(*shape), vecs, fn, scale=scale)
    got = np.frombuffer(open(fn + ".vti", "rb").read(), dtype=np.uint8)
    want = load("vti")[name]
    assert got.size == want.size and np.array_equal(got, want)


def test_write_to_vti_skips_and_warns(tmp_path):
    from pymoto_b200.domain import VoxelDomain
    from pymoto_b200.io import write_to_vti

    d = VoxelDomain(3, 2)
    with pytest.warns(UserWarning, match="neither cell- nor point-data"):
        write_to_vti(d, {"bad": np.zeros(5), "ok": np.zeros(6)}, str(tmp_path / "a.vti"))
    with pytest.warns(UserWarning, match="Nothing to write"):
        write_to_vti(d, {"bad": np.zeros(5)}, str(tmp_path / "b.vti"))
    assert not os.path.exists(tmp_path / "b.vti")
    with pytest.raises(ValueError):
        from pymoto_b200.io import WriteToVTI

        WriteToVTI.__init__(object.__new__(WriteToVTI), d, str(tmp_path / "c.vtk"))


@pytest.mark.parametrize("case", ["mma_mbb60x30", "mma_hex16x8x8"])
def test_mma_loop_with_oracle_chain_vs_reference_history(harness, case):
    """The whole design loop on the CPU: oracle chain (filter, SIMP, assembly, direct solve, scaled compliance objective
    and volume constraint) + the product's MMA update on the host harness, against the reference's MMA2007 history."""
    import scipy.sparse.linalg as spla

    import oracle
    from oracle import Grid
    from oracle.chain import cantilever
    from pymoto_b200.optimizers import mma_design_update

    g = load(case)
    if case == "mma_mbb60x30":
        gr = Grid(60, 30)
        nodes = gr.nodes3d()
        bc = np.concatenate([2 * nodes[0, :].ravel(), 2 * nodes[60, 0].ravel() + 1])
        f = np.zeros(gr.nnodes * 2)
        f[2 * nodes[0, 30].ravel() + 1] = -1.0
    else:
        gr = Grid(16, 8, 8)
        ndof, bc, f = cantilever(gr)
    flt = oracle.filter.DensityFilter(gr, 2.0)
    asm = oracle.assembly.Assembler(gr, oracle.assembly.stiffness_element(gr), bc=bc)
    n = gr.nel
    ops = HostOps(harness, n, 1)
    offset = np.full(n, 0.5)
    opt = dict(DEFAULTS, version=2007, a=np.zeros(1), c=np.full(1, 1e3), d=np.ones(1))
    x, xold1, xold2, sf, ghist, xs = np.full(n, 0.5), None, None, None, [], []
    for it in range(len(g["ghist"])):
        y = flt(x)
        K = asm(1e-9 + (1 - 1e-9) * y ** 3)
        u = spla.spsolve(K.tocsc(), f)
        c = u @ f
        sf = 100.0 / abs(c) if sf is None else sf
        gv = np.array([c * sf, (y.sum() - 0.5 * n) / (0.5 * n) * 10.0])
        dc = flt.sensitivity(asm.sensitivity(-u, u) * 3 * (1 - 1e-9) * y ** 2) * sf
        dvol = flt.sensitivity(np.full(n, 10.0 / (0.5 * n)))
        mma_design_update(ops, x, gv, [dc, dvol], offset, xold1, xold2, 0.0, 1.0, 0.1, opt)
        xold2, xold1 = xold1, x.copy()
        x = ops.x.copy()
        ghist.append(gv)
        xs.append(x)
    np.testing.assert_allclose(np.array(ghist), g["ghist"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(xs[0], g["x1"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(xs[-1], g["xlast"], rtol=0, atol=1e-5)
    assert np.mean(np.abs(offset / g["offset_last"] - 1.0) > 1e-9) < 0.01
