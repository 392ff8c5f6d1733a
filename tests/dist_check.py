"""Multi-GPU parity check of the z-slab path (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py

Every rank runs the design iteration on its slab through the public Module API; results are gathered and compared with
the CPU oracle on the whole grid: filter output and assembled values bit-exact, compliance 1e-6, dc/dx rtol 1e-6,
relative residual <= 1e-8, CG iteration count within +-1.  Exits non-zero on failure.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as ge

    if rank == 0:
        ge.build()
    dist.barrier()
    import pymoto_b200 as pmb
    from pymoto_b200 import device as dv
    from oracle import Grid
    from oracle.chain import ComplianceProblem

    peer_primitives(rank, world)
    cases = [((32, 16, 16), 4, dict(min_planes=2, min_dofs=0)),      # 2 split levels + replicated tail
             ((32, 16, 16), 4, dict(min_planes=2, min_dofs=10 ** 9)),  # only the finest level split
             ((16, 8, 8 * world), 4, dict(min_planes=2, min_dofs=0))]
    for (nx, ny, nz), min_size, kw in cases:
        gr = Grid(nx, ny, nz)
        P = ComplianceProblem(gr, kind="cantilever", tol=1e-8, min_size=min_size)
        x = np.random.default_rng(5).random(gr.nel)
        c_ref = P.response(x)
        dx_ref = P.sensitivity()

        dom = pmb.VoxelDomain(nx, ny, nz)
        mgs = pmb.solvers.auto_multigrid(dom, min_size=min_size)
        ctx = pmb.slab.init(dom, n_levels=len(mgs) + 1, **kw)
        part = ctx.part
        k0, k1 = part.planes(0)
        e0, e1 = part.elem_layers(0)
        plane, lay = (nx + 1) * (ny + 1) * 3, nx * ny
        x_loc = dv.to_device(x[e0 * lay:e1 * lay].copy())
        f_loc = dv.to_device(P.f[k0 * plane:k1 * plane].copy())

        flt = pmb.DensityFilter(dom, radius=2.0)
        simp = pmb.SIMP(1e-9, 3)
        asm = pmb.AssembleStiffness(dom, bc=P.bc)
        cg = pmb.solvers.CG(preconditioner=mgs[0], tol=1e-8)
        ls = pmb.LinSolve(hermitian=True, solver=cg)
        compl = pmb.Compliance()

        y = flt(x_loc)
        assert np.array_equal(y.cpu().numpy(), P.y[e0 * lay:e1 * lay]), "filter slab not bit-exact"
        K = asm(simp(y))
        # my rows of the oracle matrix are a contiguous run of its data array
        r0, r1 = k0 * plane, k1 * plane
        ref_rows = P.K.data[P.K.indptr[r0]:P.K.indptr[r1]]
        s_ref = 1e-9 + (1.0 - 1e-9) * P.y ** 3
        if np.array_equal(s_ref, P.s):  # SIMP on device is x*x*x, numpy uses pow: compare values only when identical
            pass
        K2 = asm(dv.to_device(P.s[e0 * lay:e1 * lay].copy()))
        assert np.array_equal(K2.data.cpu().numpy(), ref_rows), "assembled slab rows not bit-exact"
        K = asm(simp(y))
        u = ls(K, f_loc)
        c = compl(u, f_loc)
        relres = pmb.solvers.LinearSolver.residual(K, u, f_loc)
        du, _ = compl._sensitivity(1.0)
        dK, _ = ls._sensitivity(du)
        assert not ls.solver._did_solve, "adjoint must come from the LDAS database"
        ds = asm._sensitivity(dK)[0]
        dx = flt._sensitivity(simp._sensitivity(ds))
        parts = [torch.empty_like(dx) for _ in range(world)]
        dist.all_gather(parts, dx)
        dx_all = torch.cat(parts).cpu().numpy()
        c = float(c)
        err_c = abs(c - c_ref) / abs(c_ref)
        err_dx = np.abs(dx_all - dx_ref).max() / np.abs(dx_ref).max()
        if rank == 0:
            print(f"[dist_check] {nx}x{ny}x{nz} world={world} split levels={part.n_dist}/{len(mgs) + 1}: compliance {c!r} "
                  f"(oracle {c_ref!r}, rel {err_c:.2e}), relres {relres:.2e}, dc/dx err {err_dx:.2e}, CG its {cg.iterations} "
                  f"(oracle {P.cg.iterations}), halo exchanges {ctx.comm.exchanges}, all-reduces {ctx.comm.allreduces}")
        assert relres <= 1e-8, relres
        assert err_c <= 1e-6, err_c
        assert err_dx <= 1e-6, err_dx
        assert abs(cg.iterations - P.cg.iterations) <= 1
        ctx.comm.check_peer_timeouts()
        pmb.slab.reset()
    design_updates(rank, world)
    filterconv_slabs(rank, world)
    dist.barrier()
    if rank == 0:
        print("[dist_check] OK")
    dist.destroy_process_group()


def peer_primitives(rank, world):
    """The one-launch exchange steps (pmb_peer_halo_exchange / pmb_peer_allreduce) on their own: a seeded sequence of two-way
    and one-way halo exchanges of varying size with rank- and step-dependent data, eager and replayed from a CUDA graph,
    interleaved with small all-reduces; every received plane and every reduced value is checked exactly."""
    import pymoto_b200 as pmb

    nx, ny, nz = 12, 6, 4 * world
    dom = pmb.VoxelDomain(nx, ny, nz)
    ctx = pmb.slab.init(dom, n_levels=1)
    comm = ctx.comm
    if not getattr(comm, "fused", False):
        if rank == 0:
            print("[dist_check] peer primitives: one-launch exchange kernels not enabled (PMB_PEER_FUSED=0 or no symmetric memory)")
        pmb.slab.reset()
        return
    dev = torch.device("cuda", torch.cuda.current_device())
    plane = ((nx + 1) * (ny + 1) * 3) // 2  # two of these fit a mailbox slot
    own = 4 * plane
    rng = np.random.default_rng(11)  # the same sequence on every rank

    def pattern(r, step):
        return (torch.arange(own, dtype=torch.float64, device=dev) * 1e-3 + 1000.0 * r + 7.0 * step)

    buf = torch.zeros(own + 4 * plane, dtype=torch.float64, device=dev)
    off = 2 * plane

    def run(step, lower, upper, width):
        n = plane * width
        buf.fill_(-1.0)
        buf[off:off + own] = pattern(rank, step)
        comm.exchange(buf, off, own, plane, lower=lower, upper=upper, width=width)
        return n

    def check(step, lower, upper, n):
        lo, hi = buf[off - n:off], buf[off + own:off + own + n]
        if lower and rank > 0:
            assert torch.equal(lo, pattern(rank - 1, step)[own - n:]), ("lower halo", step)
        else:
            assert bool((lo == -1.0).all()), ("lower halo must stay untouched", step)
        if upper and rank < world - 1:
            assert torch.equal(hi, pattern(rank + 1, step)[:n]), ("upper halo", step)
        else:
            assert bool((hi == -1.0).all()), ("upper halo must stay untouched", step)

    nred = 0
    for step in range(120):
        lower, upper = [(True, True), (True, False), (False, True)][int(rng.integers(0, 3))]
        width = int(rng.integers(1, 3))
        n = run(step, lower, upper, width)
        check(step, lower, upper, n)
        if rng.random() < 0.5:
            k = int(rng.integers(1, 17))
            mine = torch.tensor(np.random.default_rng(1000 * step + rank).standard_normal(k), device=dev)
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            op = "sum" if rng.random() < 0.7 else "max"
            want = parts[0].clone()
            for q in parts[1:]:
                want = want + q if op == "sum" else torch.maximum(want, q)
            got = comm.allreduce_(mine.clone(), op)
            assert torch.equal(got, want), ("peer all-reduce", step, op, got, want)
            nred += 1
    assert comm.fast_exchanges == 120 and comm.fast_allreduces == nred
    # a captured sequence (two-way, one-way up, one-way down) replayed between eager exchanges
    src = pattern(rank, 0).clone()
    comm.barrier()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        buf[off:off + own] = src
        comm.exchange(buf, off, own, plane, lower=True, upper=True)
        comm.exchange(buf, off, own, plane, lower=True, upper=False, width=2)
        comm.exchange(buf, off, own, plane, lower=False, upper=True, width=2)
    for step in range(200, 212):
        src.copy_(pattern(rank, step))
        buf.fill_(-1.0)
        g.replay()
        check(step, True, True, 2 * plane)
        n = run(step + 50, True, True, 1)
        check(step + 50, True, True, n)
    comm.check_peer_timeouts()
    if rank == 0:
        print(f"[dist_check] peer primitives on {world} ranks: 120 + 12 x 4 one-launch halo exchanges (two-way / one-way, eager and "
              f"graph-replayed) and {nred} one-launch all-reduces exact")
    pmb.slab.reset()


def filterconv_slabs(rank, world):
    """FilterConv (the filter of examples/topology_optimization/ex_compliance_multigrid.py:83) on z-slabs: forward and
    backward against the oracle on the whole grid, symmetric / edge / constant z-faces, radius 2 and 3.2."""
    import pymoto_b200 as pmb
    from pymoto_b200 import device as dv
    from oracle import Grid
    from oracle.nextrows import FilterConv as OracleFilterConv

    nx, ny, nz = 9, 7, 6 * world
    gr = Grid(nx, ny, nz)
    dom = pmb.VoxelDomain(nx, ny, nz)
    ctx = pmb.slab.init(dom, n_levels=1)
    e0, e1 = ctx.part.elem_layers(0)
    lay = nx * ny
    rng = np.random.default_rng(21)
    x, dy = rng.random(gr.nel), rng.standard_normal(gr.nel)
    for kw in (dict(radius=2.0), dict(radius=3.2, zmin_bc="edge", zmax_bc=0.25), dict(radius=2.0, xmin_bc="wrap", xmax_bc="wrap", zmax_bc="edge")):
        ref = OracleFilterConv(gr, **kw)
        y_ref, dx_ref = ref(x.copy()), ref.sensitivity(dy.copy(), gr.nel)
        flt = pmb.FilterConv(dom, **kw)
        y = flt(dv.to_device(x[e0 * lay:e1 * lay].copy()))
        dx = flt._sensitivity(dv.to_device(dy[e0 * lay:e1 * lay].copy()))
        ey = np.abs(y.cpu().numpy() - y_ref[e0 * lay:e1 * lay]).max()
        ed = np.abs(dx.cpu().numpy() - dx_ref[e0 * lay:e1 * lay]).max()
        assert ey <= 1e-12 and ed <= 1e-12, ("FilterConv slab", kw, ey, ed)
    if rank == 0:
        print(f"[dist_check] FilterConv on {world} slabs: forward / backward match the oracle to 1e-12 (3 boundary sets)")
    pmb.slab.reset()


def design_updates(rank, world):
    """Three OC and three MMA design updates with the design vector distributed over the slabs (volume sum, Newton sums,
    maxima and step lengths all-reduced) against the numpy oracle on the whole grid (OC: first update 1e-6, later ones at the
    resolution of its bisection; MMA: designs 1e-5, responses 1e-6 relative)."""
    import pymoto_b200 as pmb
    from pymoto_b200 import device as dv
    from oracle import Grid
    from oracle.chain import ComplianceProblem
    from oracle.nextrows import MMAOracle, oc_update

    nx, ny, nz = 16, 8, 8 * world
    gr = Grid(nx, ny, nz)
    P = ComplianceProblem(gr, kind="cantilever", tol=1e-10, min_size=4)
    n = gr.nel
    # ---- oracle histories
    x = np.full(n, 0.5)
    oc_hist = []
    for _ in range(3):
        c = P.response(x)
        x = oc_update(x, P.sensitivity())
        oc_hist.append((c, x.copy()))
    P.u = None
    x = np.full(n, 0.5)
    mo, sf, mma_hist = MMAOracle(n, 2), None, []
    for _ in range(3):
        c = P.response(x)
        sf = 100.0 / abs(c) if sf is None else sf
        g = np.array([c * sf, (P.y.sum() - 0.5 * n) / (0.5 * n) * 10.0])
        dvol = P.filt.sensitivity(np.full(n, 10.0 / (0.5 * n)))
        x = mo.step(x, g, np.vstack([P.sensitivity() * sf, dvol]))
        mma_hist.append((g, x.copy()))

    # ---- the same loops on the slabs
    dom = pmb.VoxelDomain(nx, ny, nz)

    def network(with_volume):
        mgs = pmb.solvers.auto_multigrid(dom, min_size=4)
        ctx = pmb.slab.init(dom, n_levels=len(mgs) + 1, min_planes=2, min_dofs=0)
        k0, k1 = ctx.part.planes(0)
        e0, e1 = ctx.part.elem_layers(0)
        plane, lay = (nx + 1) * (ny + 1) * 3, nx * ny
        f = dv.to_device(P.f[k0 * plane:k1 * plane].copy())
        sx = pmb.Signal("x", state=dv.to_device(np.full((e1 - e0) * lay, 0.5)))
        with pmb.Network() as fn:
            sy = pmb.DensityFilter(dom, radius=2.0)(sx)
            ss = pmb.SIMP(1e-9, 3)(sy)
            sK = pmb.AssembleStiffness(dom, bc=P.bc)(ss)
            su = pmb.LinSolve(hermitian=True, solver=pmb.solvers.CG(preconditioner=mgs[0], tol=1e-10))(sK, f)
            sc = pmb.Compliance()(su, f)
            if with_volume:
                sg0 = pmb.Scaling(scaling=100.0)(sc)
                sg1 = pmb.Scaling(scaling=10.0, maxval=0.5 * n)(pmb.Sum()(sy))
        return sx, ([sg0, sg1] if with_volume else sc), fn, (e0 * lay, e1 * lay)

    def gathered(xloc):
        parts = [torch.empty_like(xloc) for _ in range(world)]
        dist.all_gather(parts, xloc.contiguous())
        return torch.cat(parts).cpu().numpy()

    sx, sc, fn, _ = network(False)
    oc = pmb.OC(sx, sc, fn, verbosity=0)
    xl = None
    for it, (c_ref, x_ref) in enumerate(oc_hist):
        xl, g, _ = oc.step(xl)
        xg = gathered(xl)
        # the bisection on the multiplier stops at l2 - l1 <= 1e-4 and branches on the sign of (volume - target).  Once many
        # elements sit on their move limits the volume is piecewise constant in the multiplier and can hit the target EXACTLY:
        # the branch then depends on the last bit of the sum (summation order), and the multiplier -- hence the design -- moves
        # by the width of the plateau, O(1e-4) (observed: inputs equal to 1e-13, designs 5e-5 apart, uniformly).  The first
        # update is compared tightly, the following ones at the bisection's own resolution: the loop stops at an ABSOLUTE
        # bracket width of 1e-4 on the multiplier, which on the tall 16x8x64 grid of the 8-rank run (multiplier ~ 1e-2) is a
        # relative 1e-2, i.e. designs up to ~5e-3 apart (observed 3.2e-3 with inputs equal to 1e-13); the mean density -- what the
        # bisection actually controls -- must still agree to its tolerance.
        tol_x, tol_c = (1e-6, 1e-6) if it == 0 else (5e-3, 1e-3)
        if rank == 0:
            print(f"[dist_check]   OC update {it}: objective {g!r} (oracle {c_ref!r}, rel {abs(g - c_ref) / abs(c_ref):.2e}), "
                  f"max |x - x_oracle| = {np.abs(xg - x_ref).max():.2e}, mean {np.abs(xg - x_ref).mean():.2e}")
        assert abs(g - c_ref) <= tol_c * abs(c_ref), ("OC objective", it, g, c_ref)
        assert np.abs(xg - x_ref).max() <= tol_x, ("OC design", it, np.abs(xg - x_ref).max())
        assert abs(xg.mean() - x_ref.mean()) <= 5e-4, ("OC volume", it, xg.mean(), x_ref.mean())
    if rank == 0:
        print(f"[dist_check] OC on {world} slabs: 3 updates match the oracle (first 1e-6, then the bisection tolerance), last objective {g!r}")
    pmb.slab.reset()

    sx, resp, fn, _ = network(True)
    mma = pmb.MMA(sx, resp, fn, verbosity=0)
    xl = None
    for it, (g_ref, x_ref) in enumerate(mma_hist):
        xl, g, _ = mma.step(xl)
        xg = gathered(xl)
        np.testing.assert_allclose(np.asarray(g, dtype=float), g_ref, rtol=1e-6, atol=1e-8)
        assert np.abs(xg - x_ref).max() <= 1e-5, ("MMA design", it, np.abs(xg - x_ref).max())
    if rank == 0:
        print(f"[dist_check] MMA on {world} slabs: 3 updates match the oracle (|dx| <= 1e-5), Newton iterations {mma.newton_iterations}")
    pmb.slab.reset()


if __name__ == "__main__":
    main()
