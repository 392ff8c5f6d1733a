"""Optimality-criteria update on the device.

Mirrors ``pymoto.OC`` / ``pymoto.minimize_oc`` (pymoto/common/optimizers.py:336-437, pymoto/routines.py:300-317) for
the case the compliance examples use: one design-variable Signal, one (objective) response, scalar ``move`` /
``xmin`` / ``xmax``.  The design vector and its sensitivity stay CUDA tensors; the bisection on the Lagrange multiplier
evaluates each candidate with one fused clip + deterministic-sum kernel (``pmb_oc_candidate``) and reads back one scalar.
MMA stays host code (it can drive the same Network through numpy Signals).
"""
import warnings

import numpy as np
import torch

from . import _lib
from . import device as dv


class OC:
    def __init__(self, variables, response, function, move=0.1, xmin=0.0, xmax=1.0, verbosity: int = 2, l1init: float = 0.0,
                 l2init: float = 100000.0, l1l2tol: float = 1e-4, maxvol: float = None):
        if isinstance(variables, (list, tuple)):
            if len(variables) != 1:
                raise NotImplementedError("pymoto_b200.OC handles one design-variable Signal")
            variables = variables[0]
        self.variable, self.response, self.function = variables, response, function
        self.move, self.xmin, self.xmax = float(move), float(xmin), float(xmax)
        self.dx = self.xmax - self.xmin
        self.verbosity = verbosity
        self.l1init, self.l2init, self.l1l2tol, self.maxvol = l1init, l2init, l1l2tol, maxvol
        self.iter = 0
        self._host = not dv.is_device(self.variable.state)

    def _update(self, x, dg):
        """xnew from x and dg (CUDA tensors): bisection of optimizers.py:425-435."""
        n = x.numel()
        if self.maxvol is None:
            self.maxvol = float(x.sum().item()) / n
        maxdg = float(dg.max().item())
        if maxdg > 1e-15:
            warnings.warn(f"OC only works for negative sensitivities: max(dgdx) = {maxdg}. Clipping positive values.")
        ws = dv.workspace().red
        out = dv.empty(1)
        xnew = dv.empty(n)
        l1, l2 = self.l1init, self.l2init
        lmid = 0.5 * (l1 + l2)
        target = self.maxvol * n
        while l2 - l1 > self.l1l2tol:
            lmid = 0.5 * (l1 + l2)
            _lib.call("pmb_oc_candidate", n, dv.ptr(x), dv.ptr(dg), self.move, self.xmin, self.xmax, lmid, None, dv.ptr(out),
                      dv.ptr(ws), dv.stream())
            l1, l2 = (lmid, l2) if float(out.item()) - target > 0 else (l1, lmid)
        _lib.call("pmb_oc_candidate", n, dv.ptr(x), dv.ptr(dg), self.move, self.xmin, self.xmax, lmid, dv.ptr(xnew), dv.ptr(out),
                  dv.ptr(ws), dv.stream())
        return xnew

    def step(self, x=None):
        if x is not None:
            self.variable.state = x.cpu().numpy() if self._host else x
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
            rel_step = float(torch.linalg.vector_norm((xval - xnew) / self.dx) / torch.linalg.vector_norm(xval / self.dx))
            if rel_step < tolx:
                if self.verbosity >= 1:
                    print(f"OC converged: Relative stepsize |Δx|/|x| ({rel_step}) below tolerance ({tolx})")
                break
            xval = xnew
            self.iter += 1
        # leave the network evaluated at the last accepted design (like the reference's loop: step() sets self.x)
        self.variable.state = xval.cpu().numpy() if self._host else xval
        return xval


def minimize_oc(variables, objective, function=None, maxit: int = 100, tolx: float = 1e-4, tolf: float = 1e-4, **kwargs):
    """``pymoto.minimize_oc`` with the OC update on the GPU (pymoto/routines.py:300-317)."""
    oc = OC(variables, objective, function, **kwargs)
    oc.optimize(maxiter=maxit, tolx=tolx, tolf=tolf)
    return oc
