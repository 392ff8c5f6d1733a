"""Optimiser updates on the device: OC and MMA.

MMA: see :class:`MMA` / :func:`mma_subsolv` / :class:`MmaDeviceOps` below (pymoto/common/mma.py).  OC mirrors ``pymoto.OC`` / ``pymoto.minimize_oc`` (pymoto/common/optimizers.py:336-437, pymoto/routines.py:300-317) for
the case the compliance examples use: one design-variable Signal, one (objective) response, scalar ``move`` /
``xmin`` / ``xmax``.  The design vector and its sensitivity stay CUDA tensors; the bisection on the Lagrange multiplier
evaluates each candidate with one fused clip + deterministic-sum kernel (``pmb_oc_candidate``) and reads back one scalar.
"""
import warnings

import numpy as np
import torch

from . import _lib
from . import device as dv


def _comm():
    """The slab communicator when the design vector is distributed over z-slabs (every rank holds its element layers and
    the same global scalars), else None.  Only sums / maxima over the design vector cross ranks."""
    from . import slab

    ctx = slab.context()
    return ctx.comm if ctx.active else None


def _allreduce(t, op="sum"):
    """In-place all-reduce of a small device tensor over the slabs (no-op on one GPU)."""
    comm = _comm()
    if comm is not None:
        comm.allreduce_(t, op)
    return t


def _gnorm(v):
    """Global 2-norm of a slab-distributed vector."""
    return float(torch.sqrt(_allreduce((v * v).sum().reshape(1)))[0])


def select_network(function, variables, responses, slice_network=False):
    """The Network an optimiser evaluates (pymoto/common/optimizers.py:52-72): the given one, else the outermost active
    Network; with ``slice_network`` only the modules that connect the variable Signals to the response Signals, after
    running whatever the variables themselves depend on when they have no state yet."""
    from .core import Network

    if function is None:
        if not Network.active:
            raise RuntimeError("No Network given and no active Network to take the optimisation problem from")
        function = Network.active[0]
    if not slice_network:
        return function
    subfn = function.get_output_cone(responses).get_input_cone(variables)
    if len(subfn) == 0:
        raise RuntimeError(f"Could not find a network that uses the provided input signals {variables} and produces the requested "
                           f"output signals {responses}")
    for s in variables:
        if s.state is None:
            function.get_output_cone(tosig=s).response()
        if s.state is None:
            raise RuntimeError(f"Input signal {s} has no state.")
    return subfn


class OC:
    def __init__(self, variables, response, function=None, slice_network=False, move=0.1, xmin=0.0, xmax=1.0, verbosity: int = 2, l1init: float = 0.0,
                 l2init: float = 100000.0, l1l2tol: float = 1e-4, maxvol: float = None):
        if isinstance(variables, (list, tuple)):
            if len(variables) != 1:
                raise NotImplementedError("pymoto_b200.OC handles one design-variable Signal")
            variables = variables[0]
        self.variable, self.response = variables, response
        self.function = select_network(function, [variables], [response], slice_network)
        if any(np.size(v) != 1 for v in (move, xmin, xmax)):
            raise NotImplementedError("pymoto_b200.OC takes scalar move / xmin / xmax (vector bounds: use pymoto_b200.MMA)")
        self.move, self.xmin, self.xmax = (float(np.asarray(v).reshape(-1)[0]) for v in (move, xmin, xmax))
        self.dx = self.xmax - self.xmin
        self.verbosity = verbosity
        self.l1init, self.l2init, self.l1l2tol, self.maxvol = l1init, l2init, l1l2tol, maxvol
        self.iter = 0
        self._host = not dv.is_device(self.variable.state)

    def _update(self, x, dg):
        """xnew from x and dg (CUDA tensors): bisection of optimizers.py:425-435."""
        n = x.numel()
        nglob = int(_allreduce(torch.tensor([float(n)], dtype=torch.float64, device=x.device))[0].item())
        if self.maxvol is None:
            self.maxvol = float(_allreduce(x.sum().reshape(1))[0].item()) / nglob
        maxdg = float(_allreduce(dg.max().reshape(1), "max")[0].item())
        if maxdg > 1e-15:
            warnings.warn(f"OC only works for negative sensitivities: max(dgdx) = {maxdg}. Clipping positive values.")
        ws = dv.workspace().red
        out = dv.empty(1)
        xnew = dv.empty(n)
        l1, l2 = self.l1init, self.l2init
        lmid = 0.5 * (l1 + l2)
        target = self.maxvol * nglob
        while l2 - l1 > self.l1l2tol:
            lmid = 0.5 * (l1 + l2)
            _lib.call("pmb_oc_candidate", n, dv.ptr(x), dv.ptr(dg), self.move, self.xmin, self.xmax, lmid, None, dv.ptr(out),
                      dv.ptr(ws), dv.stream())
            l1, l2 = (lmid, l2) if float(_allreduce(out).item()) - target > 0 else (l1, lmid)  # every rank takes the same branch
        _lib.call("pmb_oc_candidate", n, dv.ptr(x), dv.ptr(dg), self.move, self.xmin, self.xmax, lmid, dv.ptr(xnew), dv.ptr(out),
                  dv.ptr(ws), dv.stream())
        return xnew

    def step(self, x=None):
        if x is not None:
            self.variable.state = x.cpu().numpy() if self._host else x
            self.function.response()
        elif self.response.state is None:
            self.function.response()
        g = float(self.response.state)
        self.function.reset()
        self.response.sensitivity = 1.0
        self.function.sensitivity()
        dg = dv.to_device(self.variable.sensitivity)
        self.function.reset()
        xcur = dv.to_device(self.variable.state)
        return self._update(xcur, dg), g, dg

    def optimize(self, maxiter: int = 100, tolx: float = 1e-4, tolf: float = 1e-4):
        xval = dv.to_device(self.variable.state).clone()
        gcur = 0.0
        first = True
        while self.iter < maxiter:
            xnew, g, dg = self.step(None if first else xval)
            first = False
            gprev, gcur = gcur, g
            rel_df = abs(gcur - gprev) / abs(gcur)
            if rel_df < tolf:
                if self.verbosity >= 1:
                    print(f"OC converged: Relative function change |Δf|/|f| ({rel_df}) below tolerance ({tolf})")
                break
            if self.verbosity >= 2:
                print("It. {0: 4d}, g0({1:s}): {2:+.4e}".format(self.iter, getattr(self.response, "tag", ""), g))
            rel_step = _gnorm((xval - xnew) / self.dx) / _gnorm(xval / self.dx)
            if rel_step < tolx:
                if self.verbosity >= 1:
                    print(f"OC converged: Relative stepsize |Δx|/|x| ({rel_step}) below tolerance ({tolx})")
                break
            xval = xnew
            self.iter += 1
        # xval is the last design the loop produced; when the loop ran out of iterations it has not been evaluated yet:
        # evaluate it so that every Signal of the network belongs to variable.state (the reference leaves them consistent)
        self.variable.state = xval.cpu().numpy() if self._host else xval
        if self.iter >= maxiter and self.function is not None:
            self.function.response()
        return xval


def minimize_oc(variables, objective, function=None, maxit: int = 100, tolx: float = 1e-4, tolf: float = 1e-4, **kwargs):
    """``pymoto.minimize_oc`` with the OC update on the GPU (pymoto/routines.py:300-317)."""
    oc = OC(variables, objective, function, **kwargs)
    oc.optimize(maxiter=maxit, tolx=tolx, tolf=tolf)
    return oc


# ====================================================================================================== MMA on the device
class MmaDeviceOps:
    """The n-sized parts of one MMA subproblem as libpmb passes over CUDA tensors (``pmb_mma_*``, pmb_optim.cu).

    Holds the subproblem state (x, xsi, eta, the Newton direction, asymptotes, bounds, P, Q) in HBM; every method launches
    one fused pass and returns the handful of reduced scalars the host-side Newton driver (:func:`mma_subsolv`) needs."""

    def __init__(self, n, m):
        import ctypes as C

        dv.require_cuda()
        if not 1 <= m <= _lib.MMA_MAXM:
            raise NotImplementedError(f"pymoto_b200.MMA handles 1..{_lib.MMA_MAXM} constraints (got {m})")
        self.n, self.m = int(n), int(m)
        self.t = {nm: dv.empty(n * (m + 1) if nm in ("P", "Q") else n) for nm in _lib.MmaVecs.NAMES}
        self.vecs = _lib.MmaVecs(*[self.t[nm].data_ptr() for nm in _lib.MmaVecs.NAMES])
        self.ws = dv.zeros(_lib.query("pmb_mma_ws_doubles"))
        self.out = dv.empty(2 * _lib.MMA_MAXM + _lib.MMA_MAXM ** 2 + 8)  # the largest pass returns 2m + m^2 sums
        self._C = C
        # design vector distributed over z-slabs: n is this rank's share, the reduced sums / maxima are made global
        self.n_global = int(_allreduce(torch.tensor([float(n)], dtype=torch.float64, device=self.out.device))[0].item())

    def _reduce(self, nsum, nmax=0, first=0):
        """Make out[first : first+nsum] (sums) and the nmax entries after them (maxima) global."""
        if _comm() is not None:
            if nsum:
                _allreduce(self.out[first: first + nsum])
            if nmax:
                _allreduce(self.out[first + nsum: first + nsum + nmax], "max")

    def _host(self, vals):
        return (self._C.c_double * len(vals))(*[float(v) for v in vals])

    @staticmethod
    def _bound(b):
        if dv.is_device(b):
            return _lib.Bound(0.0, b.data_ptr())
        return _lib.Bound(float(b), None)

    @property
    def x(self):
        return self.t["x"]

    @staticmethod
    def zeros(n):
        return dv.zeros(n)

    def asymptotes(self, x, xold1, xold2, offset, asyincr, asydecr, asybound):
        _lib.call("pmb_mma_asymptotes", self.n, dv.ptr(x), dv.ptr(xold1), dv.ptr(xold2), float(asyincr), float(asydecr),
                  float(asybound), dv.ptr(offset), dv.stream())

    def setup(self, xval, dg_rows, offset, xmin, xmax, move, albefa, rho, version):
        rows = (self._C.c_void_p * (self.m + 1))(*[r.data_ptr() for r in dg_rows])
        self._keep = (dg_rows, xmin, xmax, move)  # the kernel reads them asynchronously
        _lib.call("pmb_mma_setup", self.n, self.m, dv.ptr(xval), rows, dv.ptr(offset), self._bound(xmin), self._bound(xmax),
                  self._bound(move), float(albefa), self._host(rho), int(version), self._C.byref(self.vecs), dv.ptr(self.out),
                  dv.ptr(self.ws), dv.stream())
        self._reduce(self.m + 1)
        return self.out[: self.m + 1].cpu().numpy()

    def _resid_out(self):
        self._reduce(self.m + 1, 1)
        o = self.out[: self.m + 2].cpu().numpy()
        return float(o[0]), o[1: self.m + 1].copy(), float(o[self.m + 1])

    def residual(self, lam, epsi):
        _lib.call("pmb_mma_residual", self.n, self.m, self._C.byref(self.vecs), self._host(lam), float(epsi), dv.ptr(self.out),
                  dv.ptr(self.ws), dv.stream())
        return self._resid_out()

    def newton_sums(self, lam, epsi):
        m = self.m
        _lib.call("pmb_mma_newton_sums", self.n, m, self._C.byref(self.vecs), self._host(lam), float(epsi), dv.ptr(self.out),
                  dv.ptr(self.ws), dv.stream())
        self._reduce(2 * m + m * m)
        o = self.out[: 2 * m + m * m].cpu().numpy()
        return o[:m].copy(), o[m: 2 * m].copy(), o[2 * m:].reshape(m, m).copy()

    def newton_dir(self, lam, dlam, epsi):
        _lib.call("pmb_mma_newton_dir", self.n, self.m, self._C.byref(self.vecs), self._host(lam), self._host(dlam), float(epsi),
                  dv.ptr(self.out), dv.ptr(self.ws), dv.stream())
        self._reduce(0, 4, first=1)
        return self.out[1:5].cpu().numpy()

    def linesearch(self, lam, steg, epsi):
        _lib.call("pmb_mma_linesearch", self.n, self.m, self._C.byref(self.vecs), self._host(lam), float(steg), float(epsi),
                  dv.ptr(self.out), dv.ptr(self.ws), dv.stream())
        return self._resid_out()

    def rho_sums(self, dg_rows, xmin, xmax):
        """sum_j (xmax_j - xmin_j) |dg_i[j]| for the m+1 rows (GCMMA, mma.py:151), global over slabs."""
        rows = (self._C.c_void_p * (self.m + 1))(*[r.data_ptr() for r in dg_rows])
        self._keep = (dg_rows, xmin, xmax)
        _lib.call("pmb_mma_gcmma_rho", self.n, self.m, rows, self._bound(xmin), self._bound(xmax), dv.ptr(self.out), dv.ptr(self.ws),
                  dv.stream())
        self._reduce(self.m + 1)
        return self.out[: self.m + 1].cpu().numpy()

    def estimate(self, xval, xmin, xmax):
        """Approximation values at the subproblem solution (before ``- rhs``) and dk (GCMMA, mma.py:236-239)."""
        self._keep = (xval, xmin, xmax)
        _lib.call("pmb_mma_gcmma_estimate", self.n, self.m, self._C.byref(self.vecs), dv.ptr(xval), self._bound(xmin), self._bound(xmax),
                  dv.ptr(self.out), dv.ptr(self.ws), dv.stream())
        self._reduce(self.m + 2)
        o = self.out[: self.m + 2].cpu().numpy()
        return o[: self.m + 1].copy(), float(o[self.m + 1])


def mma_subsolv(ops, m, epsimin, a0, a, b, c, d, maxittt=400):
    """Primal-dual Newton solution of the MMA subproblem, pymoto/common/mma.py:246-474, with every n-sized expression
    delegated to ``ops`` (the device) and the m-sized unknowns handled here exactly as in the reference.  The primal
    solution is left in ``ops.x``; returns (y, z, lam, mu, zet, s, number of Newton iterations)."""
    a, b, c, d = (np.asarray(v, dtype=float) for v in (a, b, c, d))
    epsi = 1.0
    y, z, lam = np.ones(m), 1.0, np.ones(m)
    mu, zet, s = np.maximum(1, 0.5 * c), 1.0, np.ones(m)
    newton_its = 0

    def small_residual(gvec):
        return np.concatenate([c + d * y - mu - lam, [a0 - zet - a @ lam], gvec - a * z - y + s - b, mu * y - epsi, [zet * z - epsi],
                               lam * s - epsi])

    while epsi > epsimin:
        sumsq, gvec, maxsq = ops.residual(lam, epsi)
        r2 = small_residual(gvec) ** 2
        residunorm, residumax = sumsq + r2.sum(), max(maxsq, r2.max())
        ittt = 0
        while residumax > (0.9 * epsi) ** 2 and ittt < maxittt:
            ittt += 1
            newton_its += 1
            gvec, GGd, GDG = ops.newton_sums(lam, epsi)
            dely = c + d * y - lam - epsi / y
            delz = a0 - a @ lam - epsi / z
            dellam = gvec - a * z - y - b + epsi / lam
            diagy = d + mu / y
            diaglamyi = s / lam + 1.0 / diagy
            AA = np.empty((m + 1, m + 1))
            bb = np.empty(m + 1)
            bb[:-1] = dellam + dely / diagy - GGd
            bb[-1] = delz
            AA[:-1, :-1] = np.diag(diaglamyi) + GDG
            AA[-1, :-1] = a
            AA[:-1, -1] = a
            AA[-1, -1] = -zet / z
            solut = np.linalg.solve(AA, bb)
            dlam, dz = solut[:m], solut[m]
            cand = ops.newton_dir(lam, dlam, epsi)  # also stores the line-search base point
            dy = -dely / diagy + dlam / diagy
            dmu = -mu + epsi / y - (mu * dy) / y
            dzet = -zet + epsi / z - zet * dz / z
            ds = -s + epsi / lam - (s * dlam) / lam
            stmxx = max(-1.01 * np.min(dy / y), -1.01 * dz / z, -1.01 * np.min(dlam / lam), 1.01 * cand[0], 1.01 * cand[1],
                        -1.01 * np.min(dmu / mu), -1.01 * dzet / zet, -1.01 * np.min(ds / s))
            steg = 1.0 / max(1.01 * cand[2], 1.01 * cand[3], stmxx, 1.0)
            yold, zold, lamold, muold, zetold, sold = y.copy(), z, lam.copy(), mu.copy(), zet, s.copy()
            for _ in range(maxittt):
                y = yold + steg * dy
                z = zold + steg * dz
                lam = lamold + steg * dlam
                mu = muold + steg * dmu
                zet = zetold + steg * dzet
                s = sold + steg * ds
                sumsq, gvec, maxsq = ops.linesearch(lam, steg, epsi)
                r2 = small_residual(gvec) ** 2
                resinorm = sumsq + r2.sum()
                if resinorm < residunorm:
                    break
                steg /= 2
            residunorm, residumax = resinorm, max(maxsq, r2.max())
        if ittt > maxittt - 2:
            print(f"MMA Subsolver: itt = {ittt}, at epsi = {'%.3e' % epsi}")
        epsi /= 10
    return y, z, lam, mu, zet, s, newton_its


def mma_subproblem(ops, x, g, dg, offset, xmin, xmax, move, opt, estimate=False):
    """Set up and solve one MMA subproblem around ``x`` (mmasub, mma.py:170-244) with the asymptote offsets as they are.
    ``opt["rho"]``: one value (MMA2007) or one per response (GCMMA; a single response's value also serves the dummy constraint
    of an unconstrained problem, as the reference's broadcast does).  The new design is left in ``ops.x``.  With ``estimate``
    the values of the convex approximations at the new design and the step measure are returned too (``gest``, ``dk``,
    :236-239).  Returns (lam, Newton iterations, gest, dk)."""
    m, n = ops.m, ops.n
    unconstrained = g.size == 1
    if unconstrained:  # dummy constraint with zero sensitivities (mma.py:172-175)
        g = np.hstack((g, -1.0))
        dg = dg + [ops.zeros(n)]
    rho = np.atleast_1d(np.asarray(opt["rho"], dtype=float))
    if rho.size == 1:
        rho = np.full(m + 1, rho[0])
    sums = ops.setup(x, dg, offset, xmin, xmax, move, opt["albefa"], list(rho), opt["version"])
    rhs = sums - g
    epsimin_scaled = opt["epsimin"] * np.sqrt(m + getattr(ops, "n_global", n))
    y, z, lam, mu, zet, s, its = mma_subsolv(ops, m, epsimin_scaled, opt["a0"], opt["a"], rhs[1:], opt["c"], opt["d"])
    gest = dk = None
    if estimate:
        est, dk = ops.estimate(x, xmin, xmax)
        gest = est - rhs
        if unconstrained:
            gest = gest[:1]
    return lam, its, gest, dk


def mma_design_update(ops, x, g, dg, offset, xold1, xold2, xmin, xmax, move, opt):
    """One MMA design update (mma.py:101-244, non-GCMMA branch): asymptote offsets from the last two designs, subproblem
    set-up, primal-dual solve.  ``ops`` owns the n-sized state (the new design is left in ``ops.x``); ``offset`` is updated
    in place; ``g`` (host, one value per response) and ``dg`` (list of rows) are not modified.  Returns (lam, Newton its)."""
    if xold1 is not None and xold2 is not None:
        ops.asymptotes(x, xold1, xold2, offset, opt["asyincr"], opt["asydecr"], opt["asybound"])
    lam, its, _, _ = mma_subproblem(ops, x, g, dg, offset, xmin, xmax, move, opt)
    return lam, its


def gcmma_rho_update(rho, g, gest, dk):
    """Raise the conservativeness of the responses whose approximation under-estimated the new value (mma.py:152-154)."""
    delta = (g - gest) / dk
    upd = delta > 0
    rho = rho.copy()
    rho[upd] = np.minimum(1.1 * (rho + delta), 10 * rho)[upd]
    return rho


def gcmma_design_update(ops, x, gk, dg, offset, xmin, xmax, move, opt, evaluate, maxit=20, verbosity=0):
    """The inner iterations of one GCMMA design update (mma.py:142-160) around the outer design ``x`` with responses ``gk``
    and sensitivity rows ``dg``; the asymptote offsets are updated by the caller beforehand.  ``evaluate(xcand)`` returns
    the responses at a candidate design (``xcand`` is ``ops.x``: copy it if it is kept).  Each pass solves the subproblem
    with P / Q of MMA2007 and ``max(rho_i, 1e-6)`` per response; the loop ends when every approximation was conservative
    (``gest >= g``) at the last candidate, which then is the new design left in ``ops.x``.
    Returns a dict: lam, newton_its, g (at the last evaluated design), rho, gest, dk, inner (subproblems solved)."""
    n_global = getattr(ops, "n_global", ops.n)
    rows = list(dg) if len(dg) > 1 else list(dg) + [ops.zeros(ops.n)]
    rho = 0.1 / n_global * ops.rho_sums(rows, xmin, xmax)[: len(dg)]
    g, gest, dk, lam, newton_its, inner = gk, None, None, None, 0, 0
    for gcmmait in range(maxit):
        if gcmmait > 0:
            g = np.asarray(evaluate(ops.x), dtype=float)
            if np.all(gest >= g):
                if verbosity >= 3:
                    print(f"  || GCMMA converged in {gcmmait} inner iterations")
                break
            rho = gcmma_rho_update(rho, g, gest, dk)
            if verbosity >= 3:
                print(f"  || GCMMA It. {gcmmait}, g = {g}, gest = {gest}, rho={rho}")
        lam, its, gest, dk = mma_subproblem(ops, x, gk, list(dg), offset, xmin, xmax, move, dict(opt, rho=np.maximum(rho, 1e-6)),
                                            estimate=True)
        newton_its += its
        inner += 1
    return dict(lam=lam, newton_its=newton_its, g=g, rho=rho, gest=gest, dk=dk, inner=inner)


class MMA:
    """``pymoto.MMA`` (pymoto/common/mma.py:5-244, base class pymoto/common/optimizers.py:9-334) with the design vector, its
    sensitivities, the asymptotes and the whole n-sized subproblem resident on the GPU.

    Same constructor and keyword options as the reference (``move, xmin, xmax, mmaversion ("MMA1987" | "MMA2007"), a0,
    epsimin, cCoef, albefa, asyinit, asyincr, asydecr, asybound, a, c``); bounds and move limits may be scalars, one value
    per variable Signal, or full vectors.  Variable states may be numpy arrays or CUDA tensors (each Signal keeps its
    kind).  ``mmaversion="GCMMA"`` runs the globally convergent inner loop (``gcmma_maxit``), ``slice_network=True`` evaluates
    only the modules between the variables and the responses."""

    def __init__(self, variables, responses, function, slice_network=False, move=0.1, xmin=0.0, xmax=1.0, verbosity=2,
                 mmaversion="MMA2007", **kwargs):
        dv.require_cuda()
        self.variables = list(variables) if isinstance(variables, (list, tuple)) else [variables]
        self.responses = list(responses) if isinstance(responses, (list, tuple)) else [responses]
        self.function, self.verbosity = select_network(function, self.variables, self.responses, slice_network), verbosity
        self._host = [not dv.is_device(s.state) for s in self.variables]
        sizes = [int(np.size(s.state)) if h else int(s.state.numel()) for s, h in zip(self.variables, self._host)]
        self._cumlens = np.concatenate([[0], np.cumsum(sizes)]).astype(int)
        self.n = int(self._cumlens[-1])
        self.response_is_uptodate = not any(s.state is None for s in self.responses)
        self.xmin = self._parse_bound(xmin, "xmin")
        self.xmax = self._parse_bound(xmax, "xmax")
        self.move = self._parse_bound(move, "move")
        self.iter = 0
        version = str(mmaversion).lower()
        self._gcmma = False
        if "1987" in version:
            self._version = 1987
        elif "2007" in version:
            self._version = 2007
        elif "gcmma" in version:  # GCMMA's P / Q are MMA2007's with rho_i = max(rho_i, 1e-6) per response (mma.py:213-216)
            self._version, self._gcmma = 2007, True
        else:
            raise ValueError('Only "MMA1987", "MMA2007", or "GCMMA" are valid options')
        self.mmaversion = mmaversion
        self.a0 = kwargs.get("a0", 1.0)
        self.epsimin = kwargs.get("epsimin", 1e-10)
        self.cCoef = kwargs.get("cCoef", 1e3)
        self.albefa = kwargs.get("albefa", 0.1)
        self.asyinit = kwargs.get("asyinit", 0.5)
        self.asyincr = kwargs.get("asyincr", 1.2)
        self.asydecr = kwargs.get("asydecr", 0.7)
        self.asybound = kwargs.get("asybound", 10.0)
        self.gcmma_maxit = kwargs.get("gcmma_maxit", 20)
        self.gest = self.dk = None
        self.gcmma_inner_iterations = 0
        self.dx = self.xmax - self.xmin  # float or CUDA tensor
        self.xold1 = self.xold2 = None
        self.offset = torch.full((self.n,), float(self.asyinit), dtype=torch.float64, device=dv.require_cuda())
        self.m = max(1, len(self.responses) - 1)
        self.a = np.asarray(kwargs.get("a", np.zeros(self.m)), dtype=float)
        if len(self.a) != self.m:
            raise RuntimeError(f"Length of the a vector ({len(self.a)}) should be equal to # constraints ({self.m}).")
        self.c = np.asarray(kwargs.get("c", np.full(self.m, self.cCoef, dtype=float)), dtype=float)
        if len(self.c) != self.m:
            raise RuntimeError(f"Length of the c vector ({len(self.c)}) should be equal to # constraints ({self.m}).")
        self.d = np.ones(self.m)
        self.ops = MmaDeviceOps(self.n, self.m)
        self.newton_iterations = 0
        self.lam = None

    # ---- bounds: scalar (kept as float), one value per Signal, or a full vector (kept as CUDA tensor)
    def _parse_bound(self, xbnd, which="bounds"):
        if dv.is_device(xbnd):
            if xbnd.numel() != self.n:
                raise RuntimeError(f"Size of {which} ({xbnd.numel()}) should be scalar or equal to number of design variables ({self.n})")
            return xbnd.to(torch.float64).reshape(-1).contiguous()
        try:
            nbnd = np.size(xbnd)
        except ValueError:  # inhomogeneous data, e.g. [[1, 2, 3], 4]
            nbnd = len(xbnd)
        if nbnd == 1:
            return float(np.asarray(xbnd).reshape(-1)[0])
        if nbnd == len(self.variables):
            bvec = np.zeros(self.n)
            for i in range(nbnd):
                bvec[self._cumlens[i]: self._cumlens[i + 1]] = xbnd[i]
        elif nbnd == self.n:
            bvec = np.asarray(xbnd, dtype=float).reshape(-1)
        else:
            raise RuntimeError(f"Size of {which} ({nbnd}) should be either:\n - scalar\n - equal to the number of variable signals "
                               f"({len(self.variables)})\n - equal to number of design variables ({self.n})")
        return dv.to_device(bvec)

    # ---- design vector <-> Signals
    @property
    def x(self):
        parts = [dv.to_device(s.state).reshape(-1) for s in self.variables]
        return parts[0].clone() if len(parts) == 1 else torch.cat(parts)

    @x.setter
    def x(self, v):
        for i, s in enumerate(self.variables):
            part = v[self._cumlens[i]: self._cumlens[i + 1]]
            cur = dv.to_device(s.state).reshape(-1)
            if not torch.equal(cur, part):
                self.response_is_uptodate = False
            s.state = part.cpu().numpy().reshape(np.shape(s.state)) if self._host[i] else part.clone()

    def calculate_g(self):
        if not self.response_is_uptodate:
            self.function.response()
            self.response_is_uptodate = True
        vals = []
        for s in self.responses:
            if s.state is None:
                raise ValueError("Response is `None` and may not yet been calculated.")
            v = s.state
            if (dv.is_device(v) and v.numel() > 1) or (not dv.is_device(v) and np.asarray(v).size > 1):
                raise TypeError("Responses for optimziation must be scalar.")
            vals.append(float(v.item()) if dv.is_device(v) else float(np.asarray(v).reshape(-1)[0]))
        return np.array(vals)

    def calculate_dg(self):
        """One back-propagation per response (optimizers.py:162-178); rows stay on the device."""
        rows = []
        self.function.reset()
        for s_out in self.responses:
            s_out.sensitivity = s_out.state * 0 + 1.0
            self.function.sensitivity()
            parts = []
            for i, v in enumerate(self.variables):
                n_i = int(self._cumlens[i + 1] - self._cumlens[i])
                parts.append(dv.zeros(n_i) if v.sensitivity is None else dv.to_device(v.sensitivity).reshape(-1))
            rows.append(parts[0].clone() if len(parts) == 1 else torch.cat(parts))
            self.function.reset()
        return rows

    def _evaluate_candidate(self, xcand):
        self.x = xcand.clone()
        return self.calculate_g()

    def step(self, x=None, g=None, dg=None):
        """One design update (mma.py:96-168).  MMA: one subproblem.  GCMMA: inner iterations at the asymptotes and sensitivities
        of the outer design ``x``, each re-evaluating the responses at the candidate (the variable Signals are left there)."""
        if x is None:
            x = self.x
        else:
            self.x = x
        if g is None:
            g = self.calculate_g()
        if dg is None:
            dg = self.calculate_dg()  # once per outer iteration
        g = np.asarray(g, dtype=float)
        opt = dict(version=self._version, albefa=self.albefa, asyincr=self.asyincr, asydecr=self.asydecr, asybound=self.asybound,
                   a0=self.a0, a=self.a, c=self.c, d=self.d, epsimin=self.epsimin)
        if self.xold1 is not None and self.xold2 is not None:
            self.ops.asymptotes(x, self.xold1, self.xold2, self.offset, self.asyincr, self.asydecr, self.asybound)
        if self._gcmma:
            r = gcmma_design_update(self.ops, x, g, list(dg), self.offset, self.xmin, self.xmax, self.move, opt, self._evaluate_candidate,
                                    self.gcmma_maxit, self.verbosity)
            self.lam, self.newton_iterations, g = r["lam"], r["newton_its"], r["g"]
            self.rho, self.gest, self.dk, self.gcmma_inner_iterations = r["rho"], r["gest"], r["dk"], r["inner"]
        else:
            self.rho = opt["rho"] = 1e-5
            self.lam, self.newton_iterations, _, _ = mma_subproblem(self.ops, x, g, list(dg), self.offset, self.xmin, self.xmax,
                                                                    self.move, opt)
        xnew = self.ops.x.clone()
        self.xold2, self.xold1 = self.xold1, x.clone()
        return xnew, g, dg

    def optimize(self, maxiter=100, tolx=1e-4, tolf=1e-4, evaluate_last=False):
        nom = type(self).__name__
        xval = self.x
        gcur = 0.0
        while self.iter < maxiter:
            xnew, g, dg = self.step(x=xval)
            gprev, gcur = gcur, g
            rel_df = np.linalg.norm(gcur - gprev) / np.linalg.norm(gcur)
            if rel_df < tolf:
                if self.verbosity >= 1:
                    print(f"{nom} converged: Relative function change |Δf|/|f| ({rel_df}) below tolerance ({tolf})")
                break
            if self.verbosity >= 2:
                msgs = ["g{0:d}({1:s}): {2:+.4e}".format(i, getattr(s, "tag", ""), g[i]) for i, s in enumerate(self.responses)]
                tag = ("[f] " if max(g[1:]) <= 0 else "[ ] ") if len(self.responses) > 1 else ""
                print("It. {0: 4d}, {1:s}{2}".format(self.iter, tag, ", ".join(msgs)))
            rel_stepsize = _gnorm((xval - xnew) / self.dx) / _gnorm(xval / self.dx)
            if rel_stepsize < tolx:
                if self.verbosity >= 1:
                    print(f"{nom} converged: Relative stepsize |Δx|/|x| ({rel_stepsize}) below tolerance ({tolx})")
                if evaluate_last:
                    self.x = xnew
                    self.calculate_g()
                break
            xval = xnew
            self.iter += 1


def minimize_mma(variables, responses, function=None, maxit: int = 100, tolx: float = 1e-4, tolf: float = 1e-4, **kwargs):
    """``pymoto.minimize_mma`` with the MMA update on the GPU (pymoto/routines.py:320-338)."""
    mma = MMA(variables, responses, function, **kwargs)
    mma.optimize(maxiter=maxit, tolx=tolx, tolf=tolf)
    return mma
