"""Golden fixtures for the optimiser rows from the UNMODIFIED reference (pyMOTO at /root/reference).

    python tests/golden/make_golden_opt.py [oc] [mma] [subsolv] [gcmma] [asm] [vti]

  ref_oc_mbb100x50.npz   10 OC iterations of the 2-D MBB 100x50 problem (BASELINE configs[0]; pym.OC.step)
  ref_mma_mbb60x30.npz   8 MMA2007 iterations, 2-D MBB 60x30, compliance objective (Scaling 100) + volume constraint (Scaling 10)
  ref_mma_hex16x8x8.npz  6 MMA2007 iterations, 3-D cantilever 16x8x8, same responses
  ref_mma_subsolv.npz    single subproblems (pym.MMA.mmasub on seeded data): m = 1, 2, unconstrained, MMA1987, vector bounds
  ref_gcmma.npz          6 GCMMA outer iterations (pym.MMA(mmaversion="GCMMA").step through a Network) of two seeded analytic problems:
                         designs, responses, rho, number of response evaluations per outer iteration
  ref_asm_multi.npz      pym.AssembleGeneral with several element matrices / bc / add_constant on seeded inputs: the assembled matrix
                         (dense) and the sensitivities of a dyad
  ref_vti.npz            bytes of VoxelDomain.write_to_vti files (2-D with vector padding, 3-D, block vectors)
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
from _refimport import import_reference  # noqa: E402
from make_golden_opt_inputs import ASM_CASES, asm_inputs, GCMMA_CASES, SUBSOLV_CASES, gcmma_problem, subsolv_inputs, vti_inputs  # noqa: E402

pym = import_reference()
assert pym is not None, "reference not found at /root/reference"

XMIN = 1e-9


def mbb2d(nx, ny):
    d = pym.VoxelDomain(nx, ny)
    bc = np.concatenate([2 * d.nodes[0, :].flatten(), 2 * d.nodes[nx, 0].flatten() + 1])
    f = np.zeros(d.nnodes * 2)
    f[2 * d.nodes[0, ny].flatten() + 1] = -1.0
    return d, bc, f


def cantilever3d(nx, ny, nz):
    d = pym.VoxelDomain(nx, ny, nz)
    bc = d.get_dofnumber(d.nodes[0, :, :].flatten(), ndof=3).flatten()
    f = np.zeros(d.nnodes * 3)
    f[3 * d.nodes[nx, :, nz // 2].flatten() + 2] = 1.0
    return d, bc, f


def network(d, bc, f, x0, with_volume):
    sx = pym.Signal("x", state=x0.copy())
    fn = pym.Network()
    with fn:
        sy = pym.DensityFilter(d, radius=2.0)(sx)
        ss = pym.MathExpression(f"{XMIN} + {1.0 - XMIN}*inp0^3")(sy)
        sK = pym.AssembleStiffness(d, bc=bc)(ss)
        su = pym.LinSolve(symmetric=True, positive_definite=True)(sK, f)
        sc = pym.EinSum("i,i->")(su, f)
        if not with_volume:
            return sx, [sc], fn
        sg0 = pym.Scaling(scaling=100.0)(sc)
        sv = pym.EinSum("i->")(sy)
        sg1 = pym.Scaling(scaling=10.0, maxval=0.5 * d.nel)(sv)
    return sx, [sg0, sg1], fn


def oc_case():
    d, bc, f = mbb2d(100, 50)
    sx, (sc,), fn = network(d, bc, f, np.full(d.nel, 0.5), with_volume=False)
    oc = pym.OC(sx, sc, fn, verbosity=0)
    hist, x = [], sx.state.copy()
    for _ in range(10):
        xnew, g, dg = oc.step(x)
        hist.append(float(np.asarray(g).reshape(-1)[0]))
        x = xnew
    oc.x = x
    final = float(np.asarray(oc.calculate_g()).reshape(-1)[0])
    print("OC history", hist[0], "->", hist[-1], "final", final)
    np.savez_compressed(os.path.join(HERE, "ref_oc_mbb100x50.npz"), history=np.array(hist), final=np.array(final), x10=x)


def mma_history(name, d, bc, f, iters):
    sx, resp, fn = network(d, bc, f, np.full(d.nel, 0.5), with_volume=True)
    mma = pym.MMA(sx, resp, fn, verbosity=0)
    ghist, xs, x = [], [], sx.state.copy()
    for _ in range(iters):
        xnew, g, dg = mma.step(x)
        ghist.append(np.array(g, dtype=float))
        x = xnew.copy()
        xs.append(x)
    print(name, "g history", [tuple(np.round(g, 6)) for g in ghist])
    np.savez_compressed(os.path.join(HERE, f"ref_{name}.npz"), ghist=np.array(ghist), x1=xs[0], x2=xs[1], xlast=xs[-1],
                        offset_last=mma.offset)


def subsolv_case():
    out = {}
    for name in SUBSOLV_CASES:
        p = subsolv_inputs(name)
        sigs = [pym.Signal("x", state=p["x"].copy())]
        resp = [pym.Signal(f"g{i}", state=float(p["g"][i])) for i in range(p["nresp"])]
        mma = pym.MMA(sigs, resp, pym.Network(), move=p["move"], xmin=p["xmin"], xmax=p["xmax"], verbosity=0, mmaversion=p["version"])
        mma.xold1, mma.xold2 = p["xold1"].copy(), p["xold2"].copy()
        xnew, _, _ = mma.step(p["x"].copy(), p["g"].copy(), p["dg"].copy())
        out[name + "_xnew"], out[name + "_offset"] = xnew, mma.offset
        out[name + "_low"], out[name + "_upp"] = mma.low, mma.upp
        print("subsolv", name, "|xnew - x|", np.linalg.norm(xnew - p["x"]))
    np.savez_compressed(os.path.join(HERE, "ref_mma_subsolv.npz"), **out)


def gcmma_case(iters=6):
    out = {}
    for name in GCMMA_CASES:
        n, x0, responses = gcmma_problem(name)
        nresp = GCMMA_CASES[name][1]
        evals = [0]

        class Analytic(pym.Module):
            def __call__(self, x):
                evals[0] += 1
                return tuple(responses(x)[0])

            def _sensitivity(self, *dg):
                J = responses(self.sig_in[0].state)[1]
                return sum(d * J[i] for i, d in enumerate(dg) if d is not None)

        sx = pym.Signal("x", state=x0.copy())
        fn = pym.Network()
        with fn:
            resp = Analytic()(sx)
        resp = list(resp) if isinstance(resp, (tuple, list)) else [resp]
        mma = pym.MMA(sx, resp, fn, verbosity=0, mmaversion="GCMMA")
        x, xs, gs, rhos, nev = sx.state.copy(), [], [], [], []
        for _ in range(iters):
            e0 = evals[0]
            xnew, g, dg = mma.step(x)
            xs.append(xnew.copy()); gs.append(np.array(g, dtype=float)); rhos.append(np.array(mma.rho, dtype=float)); nev.append(evals[0] - e0)
            x = xnew.copy()
        print("gcmma", name, "evaluations per outer iteration", nev, "g", [tuple(np.round(g, 5)) for g in gs])
        out[name + "_x"], out[name + "_g"], out[name + "_rho"], out[name + "_nev"] = np.array(xs), np.array(gs), np.array(rhos), np.array(nev)
        out[name + "_offset"] = mma.offset
    np.savez_compressed(os.path.join(HERE, "ref_gcmma.npz"), **out)


def asm_case():
    out = {}
    for name in ASM_CASES:
        p = asm_inputs(name)
        d = pym.VoxelDomain(*p["shape"])
        mod = pym.AssembleGeneral(d, p["mats"] if len(p["mats"]) > 1 else p["mats"][0], bc=p["bc"], add_constant=p["const"])
        sigs = [pym.Signal(f"x{i}", state=x.copy()) for i, x in enumerate(p["xs"])]
        sK = mod(*sigs)
        K = sK.state
        sK.sensitivity = pym.DyadicMatrix(p["u"].copy(), p["v"].copy())
        mod.sensitivity()
        out[name + "_K"] = K.toarray()
        for i, s in enumerate(sigs):
            out[f"{name}_dx{i}"] = np.asarray(s.sensitivity)
        print("asm", name, "shape", K.shape, "nnz", K.nnz, "bcdiagval", mod.bcdiagval)
    np.savez_compressed(os.path.join(HERE, "ref_asm_multi.npz"), **out)


def vti_case():
    out = {}
    for name in ("2d", "3d", "block"):
        shape, vecs, scale = vti_inputs(name)
        d = pym.VoxelDomain(*shape)
        with tempfile.TemporaryDirectory() as tmp:
            fn = os.path.join(tmp, "out.vti")
            d.write_to_vti(vecs, fn, scale=scale)
            out[name] = np.frombuffer(open(fn, "rb").read(), dtype=np.uint8)
        print("vti", name, out[name].size, "bytes")
    np.savez_compressed(os.path.join(HERE, "ref_vti.npz"), **out)


if __name__ == "__main__":
    what = sys.argv[1:] or ["oc", "mma", "subsolv", "gcmma", "asm", "vti"]
    if "oc" in what:
        oc_case()
    if "mma" in what:
        mma_history("mma_mbb60x30", *mbb2d(60, 30), iters=8)
        mma_history("mma_hex16x8x8", *cantilever3d(16, 8, 8), iters=6)
    if "subsolv" in what:
        subsolv_case()
    if "gcmma" in what:
        gcmma_case()
    if "asm" in what:
        asm_case()
    if "vti" in what:
        vti_case()
