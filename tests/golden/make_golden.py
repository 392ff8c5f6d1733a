"""Generate golden fixtures from the UNMODIFIED reference (pyMOTO v2.0.1 at /root/reference).

Run in the build container (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/ref_<case>.npz.  Large arrays are stored as SHA-256 digests of their raw bytes
(bit-exact contract) next to the small arrays that are stored in full.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from _refimport import import_reference  # noqa: E402

pym = import_reference()
assert pym is not None, "reference not found at /root/reference"


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def designs(nel, shape):
    """Three seeded designs: uniform, random, 0/1 blocks."""
    rng = np.random.default_rng(1234)
    xr = rng.random(nel)
    nx, ny, nz = shape
    e = np.arange(nel)
    i, j, k = e % nx, (e // nx) % ny, e // (nx * ny)
    xb = (((i // 2) + (j // 2) + (k // 2)) % 2).astype(float)
    return {"uniform": np.full(nel, 0.5), "random": xr, "blocks": xb}


def gmg_chain(domain, min_size):
    mgs = [pym.solvers.GeometricMultigrid(domain)]
    while True:
        sub = mgs[-1].sub_domain
        if any(n % 2 != 0 for n in sub.size) or any(sub.size < min_size):
            break
        mgs.append(pym.solvers.GeometricMultigrid(sub))
        mgs[-2].inner_level = mgs[-1]
    return mgs


def problem(kind, nx, ny, nz):
    d = pym.VoxelDomain(nx, ny, nz)
    if kind == "cantilever":
        ndof = d.dim
        bc = d.get_dofnumber(d.nodes[0, :, :].flatten(), ndof=ndof).flatten()
        f = np.zeros(d.nnodes * ndof)
        if d.dim == 3:
            f[ndof * d.nodes[nx, :, nz // 2].flatten() + 2] = 1.0
        else:
            f[ndof * d.nodes[nx, ny // 2].flatten() + 1] = 1.0
    elif kind == "heatsink":
        ndof = 1
        bc = d.nodes[0, ny // 4:(ny + 1) - ny // 4, nz // 4:(nz + 1) - nz // 4].flatten()
        f = np.zeros(d.nnodes)
        f[d.nodes[1:, :, :].flatten()] = 1.0
    elif kind == "mbb3d":
        ndof = 3
        bc = np.unique(np.concatenate([d.nodes[0, :, :].flatten() * 3, d.nodes[:, 0, :].flatten() * 3 + 1,
                                       d.nodes[nx, :, 0].flatten() * 3 + 2]))
        f = np.zeros(d.nnodes * 3)
        f[d.nodes[0, :, nz].flatten() * 3 + 2] = -1.0
    return d, ndof, np.sort(bc), f


def run_case(name, kind, nx, ny, nz, min_size, store_matrix, tol=1e-8, radius=2.0, xmin=1e-9):
    d, ndof, bc, f = problem(kind, nx, ny, nz)
    out = {"kind": kind, "shape": np.array([nx, ny, nz]), "ndof": ndof, "bc": bc, "f_nonzero": np.flatnonzero(f),
           "f_values": f[np.flatnonzero(f)], "min_size": min_size, "tol": tol, "radius": radius, "xmin": xmin}
    for dname, x0 in designs(d.nel, (nx, ny, max(nz, 1))).items():
        sx = pym.Signal("x", state=x0.copy())
        with pym.Network() as fn:
            sf = pym.DensityFilter(d, radius=radius)(sx)
            ss = pym.MathExpression(f"{xmin} + {1.0 - xmin}*inp0^3")(sf)
            if kind == "heatsink":
                asm = pym.AssemblePoisson(d, bc=bc)
            else:
                asm = pym.AssembleStiffness(d, bc=bc)
            sK = asm(ss)
            mgs = gmg_chain(d, min_size)
            cg = pym.solvers.CG(preconditioner=mgs[0], tol=tol)
            ls = pym.LinSolve(hermitian=True, solver=cg)
            su = ls(sK, f)
            sc = pym.EinSum("i,i->")(su, f)
        K = sK.state
        assert K.has_canonical_format and K.indices.dtype == np.int32
        sc.sensitivity = 1.0
        fn.sensitivity()
        res = np.linalg.norm(K @ su.state - f) / np.linalg.norm(f)
        p = f"{dname}_"
        out[p + "x"] = x0
        out[p + "y"] = sf.state
        out[p + "u"] = su.state
        out[p + "compliance"] = float(sc.state)
        out[p + "relres"] = res
        out[p + "dcdx"] = sx.sensitivity
        out[p + "data_sha256"] = digest(K.data)
        out[p + "data_sum"] = K.data.sum()
        out[p + "diag"] = K.diagonal()
        # filter backward on a fixed seed vector (DensityFilter values are unpinned by the reference's tests)
        dy = np.random.default_rng(7).standard_normal(d.nel)
        out[p + "filter_bwd_in"] = dy
        out[p + "filter_bwd"] = pym.DensityFilter(d, radius=radius)._sensitivity(dy) if dname == "random" else 0
        if store_matrix and dname == "random":
            out["indptr"], out["indices"], out["data_random"] = K.indptr, K.indices, K.data
        out["indptr_sha256"] = digest(K.indptr)
        out["indices_sha256"] = digest(K.indices)
        out["nnz"] = K.nnz
        out["n_mg"] = len(mgs)
        out["Ke"] = asm.elmat[0]
        out["bcdiagval"] = asm.bcdiagval
        print(f"{name}/{dname}: c={float(sc.state)!r} relres={res:.3e} nnz={K.nnz} mg={len(mgs)}")
    np.savez_compressed(os.path.join(HERE, f"ref_{name}.npz"), **out)


def transfer_case():
    """R restriction / prolongation of ones (reference tests/test_solvers_multigrid.py:9-91) + a Galerkin product."""
    out = {}
    for name, shape, ndof in [("2d", (8, 6, 0), 2), ("3d", (4, 6, 8), 3), ("3d1", (4, 4, 4), 1)]:
        d = pym.VoxelDomain(*shape)
        rng = np.random.default_rng(5)
        x = rng.random(d.nel)
        if ndof == 1:
            K = pym.AssemblePoisson(d)(x)
        else:
            K = pym.AssembleStiffness(d)(x)
        mg = pym.solvers.GeometricMultigrid(d)
        mg.setup_interpolation(K)
        R = mg.R
        v = rng.standard_normal(R.shape[0])
        vc = rng.standard_normal(R.shape[1])
        Ac = (R.T @ K @ R).tocsr()
        Ac.sort_indices()
        out[name + "_shape"] = np.array(shape)
        out[name + "_ndof"] = ndof
        out[name + "_x"] = x
        out[name + "_v"] = v
        out[name + "_vc"] = vc
        out[name + "_restrict"] = R.T @ v
        out[name + "_prolong"] = R @ vc
        out[name + "_restrict_ones"] = R.T @ np.ones(R.shape[0])
        out[name + "_Ac_indptr"] = Ac.indptr
        out[name + "_Ac_indices"] = Ac.indices
        out[name + "_Ac_data"] = Ac.data
        print(f"transfer {name}: R {R.shape}, Ac nnz {Ac.nnz}")
    np.savez_compressed(os.path.join(HERE, "ref_transfer.npz"), **out)


def jacobi_cg_case():
    """CG(DampedJacobi) on an assembled cantilever matrix (reference tests/test_solvers_sparse.py:276 style)."""
    d, ndof, bc, f = problem("cantilever", 6, 4, 4)
    x = np.random.default_rng(3).random(d.nel) * 0.9 + 0.1
    K = pym.AssembleStiffness(d, bc=bc)(x)
    cg = pym.solvers.CG(K, preconditioner=pym.solvers.DampedJacobi(K, w=1.0), tol=1e-10)
    fl = f.copy()
    fl[bc] = 0
    u = cg.solve(fl)
    np.savez_compressed(os.path.join(HERE, "ref_jacobi_cg.npz"), x=x, u=u, bc=bc, f=fl, shape=np.array([6, 4, 4]))
    print("jacobi cg relres", np.linalg.norm(K @ u - fl) / np.linalg.norm(fl))


FILTERCONV_CASES = [
    # name, shape, kwargs (weights given as a seed -> random asymmetric kernel of that shape)
    ("sym3d", (7, 5, 4), dict(radius=2.0)),
    ("r3_3d", (9, 6, 5), dict(radius=3.2)),
    ("sym2d", (12, 9, 0), dict(radius=2.5)),
    ("edge_wrap", (8, 6, 5), dict(radius=2.0, xmin_bc="edge", xmax_bc="wrap", ymin_bc="wrap", ymax_bc="wrap", zmin_bc="edge", zmax_bc="edge")),
    ("const", (6, 7, 5), dict(radius=2.0, xmin_bc=0.0, xmax_bc=1.0, ymin_bc=0.25, ymax_bc="symmetric", zmin_bc="edge", zmax_bc=0.75)),
    ("weights", (6, 5, 4), dict(weights_shape=(3, 5, 3), xmin_bc="wrap", xmax_bc="symmetric", ymin_bc=0.5, ymax_bc="edge")),
    ("weights2d", (9, 8, 0), dict(weights_shape=(5, 3), xmin_bc="edge", ymax_bc=2.0)),
    ("override", (6, 6, 4), dict(radius=2.0, override=True)),
]


def filterconv_case():
    """FilterConv forward / backward for every boundary mode (reference tests/test_filter.py:13-223 pin these by impulse
    responses and finite differences; here the reference's outputs themselves are stored)."""
    out = {}
    for name, shape, kw in FILTERCONV_CASES:
        kw = dict(kw)
        d = pym.VoxelDomain(*shape)
        rng = np.random.default_rng(len(name) * 7 + shape[0])
        wshape = kw.pop("weights_shape", None)
        override = kw.pop("override", False)
        if wshape is not None:
            kw["weights"] = np.random.default_rng(len(name)).random(wshape)
            out[name + "_weights"] = kw["weights"]
        m = pym.FilterConv(d, **kw)
        if override:
            m.override_values((np.s_[1:3], np.s_[2:4], np.s_[:]), 1.0)  # a fixed solid region inside the domain
        x = rng.random(d.nel)
        dy = rng.standard_normal(d.nel)
        sx = pym.Signal("x", state=x)
        sy = m(sx)
        sy.sensitivity = dy
        m.sensitivity()
        out[name + "_x"], out[name + "_y"], out[name + "_dy"], out[name + "_dx"] = x, sy.state, dy, sx.sensitivity
        print(f"filterconv {name}: weights {m.weights.shape}, |y| {np.linalg.norm(sy.state):.6f}")
    np.savez_compressed(os.path.join(HERE, "ref_filterconv.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "filterconv":
        filterconv_case()
        sys.exit(0)
    run_case("hex_6x4x4", "cantilever", 6, 4, 4, min_size=8, store_matrix=True)
    run_case("quad_12x8", "cantilever", 12, 8, 0, min_size=4, store_matrix=True)
    run_case("hex_16x8x8", "cantilever", 16, 8, 8, min_size=4, store_matrix=False)
    run_case("thermal_8x8x8", "heatsink", 8, 8, 8, min_size=4, store_matrix=True)
    run_case("mbb_8x4x4", "mbb3d", 8, 4, 4, min_size=4, store_matrix=False)
    transfer_case()
    jacobi_cg_case()
    filterconv_case()
