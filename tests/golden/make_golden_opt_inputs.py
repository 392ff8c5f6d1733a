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
    "m5": (350, 6, "MMA2007", False),
    "m6_vecbounds": (280, 7, "MMA2007", True),
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

ASM_CASES = {
    # name: (shape, dofs per node, number of element matrices, Dirichlet dofs?, add_constant?)
    "hex_two_const": ((5, 4, 3), 3, 2, True, True),
    "quad_const": ((7, 5, 0), 1, 1, False, True),
    "quad_three_bc": ((6, 5, 0), 2, 3, True, False),
    "hex_two_thermal": ((4, 4, 5), 1, 2, True, False),
}


def asm_inputs(name):
    """Seeded inputs of an AssembleGeneral with several element matrices and / or a constant matrix (assembly.py:38-47,
    245-253, 294-295): element matrices, scaling vectors, Dirichlet dofs, the constant (scipy CSR inside the node stencil:
    a random diagonal plus couplings between the dofs of a node and to the next node in x), and a dyad (u, v) for the sensitivity."""
    import scipy.sparse as sps

    shape, ndof, nmat, with_bc, with_const = ASM_CASES[name]
    rng = np.random.default_rng(sum(map(ord, name)))
    nx, ny, nz = shape
    dim = 3 if nz > 0 else 2
    nel = nx * ny * max(nz, 1)
    nnodes = (nx + 1) * (ny + 1) * (nz + 1)
    n = nnodes * ndof
    ld = (2 ** dim) * ndof
    mats = []
    for _ in range(nmat):
        r = rng.standard_normal((ld, ld))
        mats.append(r @ r.T / ld + 0.1 * rng.standard_normal((ld, ld)))  # general (not symmetric) real matrices
    xs = [0.05 + rng.random(nel) for _ in range(nmat)]
    bc = np.unique(rng.integers(0, n, max(3, n // 12))) if with_bc else None
    const = None
    if with_const:
        diag = sps.diags(0.5 + rng.random(n))
        r = np.arange(n - ndof)
        nxt = sps.coo_matrix((0.1 * rng.standard_normal(r.size), (r, r + ndof)), shape=(n, n))  # next node in x (or wraps to the
        keep = ((r // ndof) % (nx + 1)) != nx                                                   # next row: dropped)
        nxt = sps.coo_matrix((nxt.data[keep], (nxt.row[keep], nxt.col[keep])), shape=(n, n))
        const = (diag + nxt + 0.5 * nxt.T).tocsr()
        if ndof > 1:
            rr = np.arange(n - 1)
            same = (rr // ndof) == ((rr + 1) // ndof)
            const = (const + sps.coo_matrix((0.2 * rng.standard_normal(int(same.sum())), (rr[same], rr[same] + 1)), shape=(n, n))).tocsr()
    u, v = rng.standard_normal(n), rng.standard_normal(n)
    return dict(shape=shape, ndof=ndof, mats=mats, xs=xs, bc=bc, const=const, u=u, v=v)


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
