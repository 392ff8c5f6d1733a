"""Seeded inputs of the optimiser / VTI fixtures, shared by the generator (make_golden_opt.py, needs the reference) and
the tests (which must run where the reference is absent)."""
import numpy as np

SUBSOLV_CASES = {
    # name: (n, number of responses, version, vector bounds?)
    "m1": (500, 2, "MMA2007", False),
    "m2": (300, 3, "MMA2007", False),
    "unconstrained": (200, 1, "MMA2007", False),
    "m1_1987": (400, 2, "MMA1987", False),
    "m3_vecbounds": (250, 4, "MMA2007", True),
}


def subsolv_inputs(name):
    """Seeded inputs of one MMA subproblem; shared with the tests (they import this function)."""
    n, nresp, version, vec = SUBSOLV_CASES[name]
    rng = np.random.default_rng(sum(map(ord, name)))
    x = 0.1 + 0.8 * rng.random(n)
    xold1 = np.clip(x + 0.05 * rng.standard_normal(n), 0, 1)
    xold2 = np.clip(xold1 + 0.05 * rng.standard_normal(n), 0, 1)
    g = np.concatenate([[10.0 * rng.random()], 0.3 * rng.standard_normal(nresp - 1)])
    dg = rng.standard_normal((nresp, n))
    dg[0] = -np.abs(dg[0]) * 3.0  # compliance-like objective
    if nresp > 1:
        dg[1] = 0.02 + 0.01 * rng.random(n)  # volume-like constraint
    if vec:
        xmin, xmax, move = 0.05 * rng.random(n), 1.0 - 0.05 * rng.random(n), 0.05 + 0.1 * rng.random(n)
    else:
        xmin, xmax, move = 0.0, 1.0, 0.1
    return dict(n=n, nresp=nresp, version=version, x=x, xold1=xold1, xold2=xold2, g=g, dg=dg, xmin=xmin, xmax=xmax, move=move)


GCMMA_CASES = {
    # name: (n, number of responses)
    "gcmma_m2": (150, 3),
    "gcmma_unconstrained": (90, 1),
}


def gcmma_problem(name):
    """A seeded analytic, deliberately non-convex problem (the oscillating term makes MMA approximations non-conservative, so
    GCMMA's inner iterations trigger): returns (n, x0, responses(x) -> (g[nresp], dg[nresp, n]))."""
    n, nresp = GCMMA_CASES[name]
    rng = np.random.default_rng(sum(map(ord, name)))
    w, v, u = 1.0 + rng.random(n), rng.standard_normal(n), 0.5 + rng.random(n)
    x0 = np.full(n, 0.45)

    def responses(x):
        g = [np.sum(w / (x + 0.05)) / n + 2.0 * np.sum(v * np.sin(9.0 * x)) / n]
        dg = [-w / (x + 0.05) ** 2 / n + 18.0 * v * np.cos(9.0 * x) / n]
        if nresp > 1:
            g.append(np.sum(x) / n - 0.4)
            dg.append(np.full(n, 1.0 / n))
        if nresp > 2:
            g.append(np.sum(u * (x - 0.3) ** 2 * np.cos(5.0 * x)) / n - 0.02)
            dg.append(u * (2.0 * (x - 0.3) * np.cos(5.0 * x) - 5.0 * (x - 0.3) ** 2 * np.sin(5.0 * x)) / n)
        return np.array(g), np.vstack(dg)

    return n, x0, responses


def vti_inputs(name):
    rng = np.random.default_rng(len(name))
    if name == "2d":
        shape = (5, 4, 0)
        nel, nn = 20, 30
        return shape, {"rho": rng.random(nel), "u": rng.standard_normal(2 * nn), "T": rng.standard_normal(nn)}, 1.0
    if name == "3d":
        shape = (4, 3, 2)
        nel, nn = 24, 60
        return shape, {"x": rng.random(nel), "disp": rng.standard_normal(3 * nn), "f": rng.standard_normal(3 * nn)}, 0.5
    shape = (3, 2, 2)
    nel, nn = 12, 36
    return shape, {"modes": rng.standard_normal((2, 3 * nn)), "sens": rng.standard_normal((3, nel)), "c": rng.standard_normal(nn) + 1j * rng.standard_normal(nn)}, 1.0
