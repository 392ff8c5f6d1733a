"""Oracle restatements for the "next" rows that are built: FilterConv and the OC update
(test infrastructure only).

  FilterConv           pymoto/modules/filter.py:8-220   (padded index array + scipy.signal convolve / correlate)
  OC.step (update)     pymoto/common/optimizers.py:402-437
"""
from numbers import Number

import numpy as np
from scipy.signal import convolve, correlate

from .grid import Grid


class FilterConv:
    def __init__(self, grid: Grid, radius=None, weights=None, xmin_bc="symmetric", xmax_bc="symmetric", ymin_bc="symmetric",
                 ymax_bc="symmetric", zmin_bc="symmetric", zmax_bc="symmetric"):
        self.grid = grid
        if (weights is None) == (radius is None):
            raise ValueError("Only one of arguments 'filter_radius' or 'weights' must be provided.")
        if weights is not None:
            self.weights = np.array(weights, dtype=float)
            while self.weights.ndim < 3:
                self.weights = np.expand_dims(self.weights, axis=-1)
        else:  # filter.py:189-205 (relative units)
            n = [grid.nelx, grid.nely, grid.nelz]
            d = [min(n[a], int(radius - 1e-10)) for a in range(3)]
            cx, cy, cz = np.meshgrid(*[np.arange(-v, v + 1) * 1.0 for v in d], indexing="ij")
            self.weights = np.maximum(0.0, radius - np.sqrt(cx * cx + cy * cy + cz * cz))
            self.weights /= np.sum(self.weights)
        self.pad = [v // 2 for v in self.weights.shape]
        self.n = [grid.nelx, grid.nely, max(grid.nelz, 1)]
        self.overrides = []
        ex, ey, ez = np.meshgrid(*[np.arange(v) for v in self.n], indexing="ij")
        self.el3d_orig = grid.elem_number(ex, ey, ez)
        a = self._pad(self.el3d_orig, xmin_bc, xmax_bc, 0)
        a = self._pad(a, ymin_bc, ymax_bc, 1)
        self.el3d_pad = self._pad(a, zmin_bc, zmax_bc, 2)

    def _pad(self, idx, bc0, bc1, axis):
        p = self.pad[axis]
        pw = lambda lo, hi: [(lo, hi) if a == axis else (0, 0) for a in range(3)]  # noqa: E731
        if bc0 == "wrap" or bc1 == "wrap":
            idx = np.pad(idx, pw(p if bc0 == "wrap" else 0, p if bc1 == "wrap" else 0), mode="wrap")
        padded = [self.n[a] + 2 * self.pad[a] for a in range(3)]
        for side, bc in ((1, bc1), (0, bc0)):  # the reference handles the max edge first
            width = pw(0, p) if side else pw(p, 0)
            if bc in ("edge", "symmetric"):
                idx = np.pad(idx, width, mode=bc)
            elif isinstance(bc, Number):
                idx = np.pad(idx, width, mode="constant", constant_values=0)
                rng = [np.arange(v) for v in padded]
                rng[axis] = (p + self.n[axis] + np.arange(p)) if side else np.arange(p)
                if p > 0:
                    self.overrides.append((tuple(np.meshgrid(*rng, indexing="ij")), bc))
        return idx

    def override_values(self, index, value):
        rng = [self.pad[a] + np.arange(self.n[a]) for a in range(3)]
        ex, ey, ez = np.meshgrid(*rng, indexing="ij")
        self.overrides.append(((ex[index], ey[index], ez[index]), value))

    def __call__(self, x):
        xpad = x[self.el3d_pad]
        for index, value in self.overrides:
            xpad[index] = value
        y = np.zeros_like(x)
        np.add.at(y, self.el3d_orig, convolve(xpad, self.weights, mode="valid"))
        return y

    def sensitivity(self, dy, nel):
        dx3d = correlate(dy[self.el3d_orig], self.weights, mode="full")
        for index, _ in self.overrides:
            dx3d[index] = 0
        dx = np.zeros(nel)
        np.add.at(dx, self.el3d_pad, dx3d)
        return dx


def oc_update(x, dg, move=0.1, xmin=0.0, xmax=1.0, maxvol=None, l1=0.0, l2=100000.0, tol=1e-4):
    """One optimality-criteria design update (optimizers.py:416-435)."""
    if maxvol is None:
        maxvol = np.sum(x) / x.size
    dg = np.minimum(dg, 0)
    lb, ub = np.maximum(xmin, x - move), np.minimum(xmax, x + move)
    xnew = x.copy()
    while l2 - l1 > tol:
        lmid = 0.5 * (l1 + l2)
        xnew[:] = np.clip(x * np.sqrt(-dg / lmid), lb, ub)
        l1, l2 = (lmid, l2) if np.sum(xnew) - maxvol * x.size > 0 else (l1, lmid)
    return xnew


# ------------------------------------------------------------------------------------------------------------------ MMA
class MMAOracle:
    """numpy restatement of the reference's MMA design update (pymoto/common/mma.py; test infrastructure only).

      asymptote offsets   mma.py:120-140    mmasub (bounds, P, Q, rhs)   mma.py:170-244
      subsolv             mma.py:246-474    (primal-dual Newton with the (m+1)x(m+1) reduced system and a residual line search)

    Supports MMA1987 / MMA2007 / GCMMA, scalar or vector xmin / xmax / move; ``step(x, g, dg)`` takes the responses
    g (first = objective) and their sensitivities dg (rows) and returns the new design.  GCMMA (mma.py:104-160, 232-242) needs the
    responses at the inner candidates: pass ``evaluate(x) -> g``; ``step`` then returns the design accepted by the inner loop and
    leaves the responses evaluated last in ``self.g_last`` (the reference returns them), ``self.rho``, ``self.inner``."""

    def __init__(self, n, nresp, move=0.1, xmin=0.0, xmax=1.0, version="MMA2007", a0=1.0, epsimin=1e-10, ccoef=1e3, albefa=0.1,
                 asyinit=0.5, asyincr=1.2, asydecr=0.7, asybound=10.0):
        self.n, self.m = n, max(1, nresp - 1)
        self.move, self.xmin, self.xmax = move, xmin * np.ones(n), xmax * np.ones(n)
        self.dx = self.xmax - self.xmin
        self.version = version
        self.a0, self.epsimin, self.albefa = a0, epsimin, albefa
        self.asyincr, self.asydecr, self.asybound = asyincr, asydecr, asybound
        self.offset = asyinit * np.ones(n)
        self.a, self.c, self.d = np.zeros(self.m), np.full(self.m, float(ccoef)), np.ones(self.m)
        self.xold1 = self.xold2 = None
        self.newton_iterations = 0

    def step(self, x, g, dg, evaluate=None, gcmma_maxit=20):
        g, dg = np.atleast_1d(np.asarray(g, dtype=float)), np.atleast_2d(np.asarray(dg, dtype=float))
        if self.xold1 is not None and self.xold2 is not None:  # :120-140
            zzz = (x - self.xold1) * (self.xold1 - self.xold2)
            self.offset[zzz > 0] *= self.asyincr
            self.offset[zzz < 0] *= self.asydecr
            self.offset = np.clip(self.offset, 1 / self.asybound ** 2, self.asybound)
        if "gcmma" not in self.version.lower():
            xnew = self._mmasub(x, g, dg, rho=1e-5)
        else:  # :142-160 inner iterations around the outer design x with its responses gk = g and sensitivities dg
            gk, xnew, self.inner = g, None, 0
            for it in range(gcmma_maxit):
                if it > 0:
                    g = np.atleast_1d(np.asarray(evaluate(xnew), dtype=float))
                    if np.all(self.gest >= g):
                        break
                    delta = (g - self.gest) / self.dk  # :152-154
                    upd = delta > 0
                    self.rho[upd] = np.minimum(1.1 * (self.rho + delta), 10 * self.rho)[upd]
                else:
                    self.rho = 0.1 / self.n * np.sum(self.dx * np.abs(dg), axis=1)  # :151
                xnew = self._mmasub(x, gk, dg, rho=self.rho)
                self.inner += 1
            self.g_last = g
        self.xold2, self.xold1 = self.xold1, x.copy()
        return xnew

    def _mmasub(self, xval, g, dg, rho):
        unconstrained = g.size == 1
        if unconstrained:  # :172-175 dummy constraint
            g, dg = np.hstack((g, -1.0)), np.vstack((dg, np.zeros(self.n)))
        shift = self.offset * self.dx
        self.low, self.upp = xval - shift, xval + shift
        alfa = np.maximum.reduce([self.low + self.albefa * shift, xval - self.move * self.dx, self.xmin])
        beta = np.minimum.reduce([self.upp - self.albefa * shift, xval + self.move * self.dx, self.xmax])
        gp, gm, dx2 = np.maximum(dg, 0), np.maximum(-dg, 0), shift ** 2
        if "1987" in self.version:
            P, Q = dx2 * gp, dx2 * gm
        elif "gcmma" in self.version.lower():  # :213-216 (a single response's rho also serves the dummy row: the reference broadcasts)
            rr = np.maximum(np.asarray(rho, dtype=float).reshape(-1, 1), 1e-6)
            P = dx2 * (1.001 * gp + 0.001 * gm + rr / self.dx)
            Q = dx2 * (0.001 * gp + 1.001 * gm + rr / self.dx)
        else:
            P = dx2 * (1.001 * gp + 0.001 * gm + rho / self.dx)
            Q = dx2 * (0.001 * gp + 1.001 * gm + rho / self.dx)
        rhs = P @ (1 / shift) + Q @ (1 / shift) - g
        xmma = self._subsolv(self.epsimin * np.sqrt(self.m + self.n), self.low, self.upp, alfa, beta, P, Q, rhs[1:], xval)
        # :236-239 values of the approximations at the subproblem solution and the step measure of the rho update
        self.gest = np.sum(P / (self.upp - xmma) + Q / (xmma - self.low), axis=1) - rhs
        if unconstrained:
            self.gest = self.gest[[0]]
        self.dk = np.sum((self.upp - self.low) * (xmma - xval) ** 2 / ((self.upp - xmma) * (xmma - self.low) * self.dx))
        return xmma

    def _subsolv(self, epsimin, low, upp, alfa, beta, P, Q, b, x0):
        m, a0, a, c, d = self.m, self.a0, self.a, self.c, self.d
        P0, Q0, P1, Q1 = P[0], Q[0], P[1:], Q[1:]
        x = np.clip(x0, alfa + 1e-10, beta - 1e-10)
        y, z, lam, s = np.ones(m), 1.0, np.ones(m), np.ones(m)
        xsi, eta = np.maximum(1.0 / (x - alfa), 1), np.maximum(1.0 / (beta - x), 1)
        mu, zet = np.maximum(1, 0.5 * c), 1.0

        def residual(x, y, z, lam, xsi, eta, mu, zet, s, epsi):
            ux1, xl1 = upp - x, x - low
            plam, qlam = P0 + lam @ P1, Q0 + lam @ Q1
            gvec = P1 @ (1 / ux1) + Q1 @ (1 / xl1)
            dpsidx = plam / ux1 ** 2 - qlam / xl1 ** 2
            return np.concatenate([dpsidx - xsi + eta, c + d * y - mu - lam, [a0 - zet - a @ lam], gvec - a * z - y + s - b,
                                   xsi * (x - alfa) - epsi, eta * (beta - x) - epsi, mu * y - epsi, [zet * z - epsi], lam * s - epsi])

        epsi, self.newton_iterations = 1.0, 0
        while epsi > epsimin:
            r2 = residual(x, y, z, lam, xsi, eta, mu, zet, s, epsi) ** 2
            rnorm, rmax, it = r2.sum(), r2.max(), 0
            while rmax > (0.9 * epsi) ** 2 and it < 400:
                it += 1
                self.newton_iterations += 1
                ux1, xl1 = upp - x, x - low
                plam, qlam = P0 + lam @ P1, Q0 + lam @ Q1
                gvec = P1 @ (1 / ux1) + Q1 @ (1 / xl1)
                GG = P1 / ux1 ** 2 - Q1 / xl1 ** 2
                delx = plam / ux1 ** 2 - qlam / xl1 ** 2 - epsi / (x - alfa) + epsi / (beta - x)
                dely, delz = c + d * y - lam - epsi / y, a0 - a @ lam - epsi / z
                dellam = gvec - a * z - y - b + epsi / lam
                diagx = 2 * (plam / ux1 ** 3 + qlam / xl1 ** 3) + xsi / (x - alfa) + eta / (beta - x)
                diagy = d + mu / y
                AA = np.zeros((m + 1, m + 1))
                AA[:m, :m] = np.diag(s / lam + 1.0 / diagy) + (GG / diagx) @ GG.T
                AA[m, :m] = AA[:m, m] = a
                AA[m, m] = -zet / z
                sol = np.linalg.solve(AA, np.concatenate([dellam + dely / diagy - GG @ (delx / diagx), [delz]]))
                dlam, dz = sol[:m], sol[m]
                dx = -delx / diagx - (dlam @ GG) / diagx
                dy = -dely / diagy + dlam / diagy
                dxsi = -xsi + epsi / (x - alfa) - xsi * dx / (x - alfa)
                deta = -eta + epsi / (beta - x) + eta * dx / (beta - x)
                dmu, dzet, ds = -mu + epsi / y - mu * dy / y, -zet + epsi / z - zet * dz / z, -s + epsi / lam - s * dlam / lam
                stmxx = -1.01 * min(np.min(dy / y), dz / z, np.min(dlam / lam), np.min(dxsi / xsi), np.min(deta / eta),
                                    np.min(dmu / mu), dzet / zet, np.min(ds / s))
                steg = 1.0 / max(-1.01 * np.min(dx / (x - alfa)), 1.01 * np.max(dx / (beta - x)), stmxx, 1.0)
                old = (x, y, z, lam, xsi, eta, mu, zet, s)
                step = (dx, dy, dz, dlam, dxsi, deta, dmu, dzet, ds)
                for _ in range(400):
                    x, y, z, lam, xsi, eta, mu, zet, s = (o + steg * dv_ for o, dv_ in zip(old, step))
                    r2 = residual(x, y, z, lam, xsi, eta, mu, zet, s, epsi) ** 2
                    if r2.sum() < rnorm:
                        break
                    steg /= 2
                rnorm, rmax = r2.sum(), r2.max()
            epsi /= 10
        self.lam = lam
        return x
