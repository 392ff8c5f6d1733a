#!/usr/bin/env python
"""bench.py -- design iterations / s of the compliance hot path (BASELINE.json metric).

A "step" is ONE design iteration of the 3-D cantilever compliance problem on a hex8 grid (default 256x128x128,
12.8 M dof = BASELINE configs[2]): density filter -> SIMP -> stiffness assembly -> LinSolve (LDAS + CG(tol 1e-8) +
geometric multigrid, warm-started) -> compliance -> adjoint (LDAS, no CG) -> element sensitivities -> SIMP' -> filter^T.
Between steps the design is perturbed, x <- clip(x + 0.2 (rand - 0.5), 0, 1) (tests/bench_assembly.py:80-90 of the
reference), all designs pre-generated from a fixed seed.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size NX NY NZ] [--impl b200|reference] [--problem ...]

`value`        device-resident throughput (inputs already in HBM), CUDA-event timed, max over ranks
`e2e`          the same step driven from pinned HOST buffers: x copied host->device and (compliance, dc/dx) copied
               device->host inside the timed region, through the public Module API
`roofline`     the kernel that dominates the timed step -- the matrix-free finest-level operator (FP64 bound): achieved
               TFLOP/s against the FP64 rate this GPU sustains in the same run (pmb_probe_fp64), plus its HBM view; the
               HBM-bound stencil-CSR kernel on level 1 (`csr_level1`, in the step) and level 0 (`csr_level0`, only in the
               `csr_streamed` leg) and the whole CSR-streamed iteration against SURVEY 8d's byte model (`iteration`)
`parity`       a small-grid gate against the CPU oracle run BEFORE timing (slab-decomposed when N > 1)
`cpu_baseline` the UNMODIFIED reference (pyMOTO from baseline/_ref or /root/reference) on a bounded sample grid
`secondary`    BASELINE configs[3] (3-D MBB bc set on the same grid, N > 1) and configs[4] (thermal 256^3, N = 1 and 8)
--impl reference runs only the CPU arm: the unmodified reference's own modules on the host cores (rank 0 only).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

XMIN, RADIUS, TOL = 1e-9, 2.0, 1e-8
METRIC, UNIT = "design_iters_per_sec", "iter/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, nargs=3, default=None, metavar=("NX", "NY", "NZ"))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-size", type=int, nargs=3, default=None, help="sample grid of the CPU arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the configs[3] / configs[4] side measurements")
    ap.add_argument("--no-parity", action="store_true", help="skip the small-grid oracle gate")
    ap.add_argument("--problem", default="cantilever", choices=["cantilever", "mbb", "thermal"],
                    help="cantilever = BASELINE configs[2] (metric); mbb = configs[3] bc set; thermal = configs[4]")
    ap.add_argument("--csr", action="store_true", help="stream the assembled CSR values on every level (no matrix-free level 0)")
    ap.add_argument("--kernel-only", action="store_true", help="only the dominant-kernel loop (for ncu captures)")
    ap.add_argument("--profile", action="store_true", help="per-entry-point CUDA-event breakdown of one step (diagnostic)")
    return ap.parse_args()


def design_sequence(nel, count, seed=1234, keep=None):
    """x_0 = 0.5, then successive bounded random perturbations (seeded).  ``keep`` maps a global design to the part a
    rank stores (its slab); the random stream is always the global one so every rank count sees the same designs."""
    rng = np.random.default_rng(seed)
    keep = (lambda a: a) if keep is None else keep
    x = np.full(nel, 0.5)
    out = [keep(x).copy()]
    for _ in range(count - 1):
        x = np.clip(x + 0.2 * (rng.random(nel) - 0.5), 0.0, 1.0)
        out.append(keep(x).copy())
    return out


# weak scaling: the per-GPU slab stays 256x128x128-equivalent (12.8 M dof) as ranks are added
WEAK_GRIDS = {1: (256, 128, 128), 2: (256, 128, 256), 4: (256, 256, 256), 8: (512, 256, 256)}


def problem_setup(problem, nx, ny, nz, k0=0, k1=None):
    """(ndof, sorted GLOBAL bc dofs, load on node planes [k0, k1)) of the three synthetic problems."""
    k1 = nz + 1 if k1 is None else k1
    NX, NY = nx + 1, ny + 1
    kk, jj, ii = np.arange(nz + 1), np.arange(NY), np.arange(NX)
    node = lambda i, j, k: ((np.asarray(k)[..., None, None] * NY + np.asarray(j)[..., None]) * NX + np.asarray(i)).ravel()  # noqa: E731
    if problem == "cantilever":  # all dofs clamped on face i = 0, unit +z load on the line i = nx, k = nz / 2
        ndof = 3
        face = node(np.array([0]), jj, kk)
        bc = (face[:, None] * 3 + np.arange(3)[None, :]).ravel()
        f = np.zeros((k1 - k0) * NX * NY * 3)
        kl = nz // 2
        if k0 <= kl < k1:
            f[(((kl - k0) * NY + jj) * NX + nx) * 3 + 2] = 1.0
    elif problem == "mbb":  # half-MBB: u_x = 0 on face i = 0, u_y = 0 on face j = 0, u_z = 0 on edge (i = nx, k = 0), -z line load on edge (i = 0, k = nz)
        ndof = 3
        bc = np.concatenate([node(np.array([0]), jj, kk) * 3 + 0, node(ii, np.array([0]), kk) * 3 + 1,
                             node(np.array([nx]), jj, np.array([0])) * 3 + 2])
        f = np.zeros((k1 - k0) * NX * NY * 3)
        if k0 <= nz < k1:
            f[(((nz - k0) * NY + jj) * NX + 0) * 3 + 2] = -1.0
    else:  # heat sink: T = 0 on a centred patch of face i = 0, unit heat load on every node with i >= 1
        ndof = 1
        bc = node(np.array([0]), np.arange(ny // 4, NY - ny // 4), np.arange(nz // 4, (nz + 1) - nz // 4))
        f = np.ones((k1 - k0, NY, NX))
        f[:, :, 0] = 0.0
        f = f.ravel()
    return ndof, np.unique(bc), f


# ------------------------------------------------------------------------------------------------ CPU arm: the unmodified reference
def load_reference():
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import refload

    return refload.import_reference(), refload.reference_root()


class ReferenceChain:
    """The same design iteration built from the UNMODIFIED reference's own modules (pymoto.DensityFilter, MathExpression,
    AssembleStiffness, LinSolve(CG(GeometricMultigrid chain)), EinSum) -- its stock scipy code path, one Network."""

    def __init__(self, pym, size, problem="cantilever"):
        nx, ny, nz = size
        self.pym = pym
        dom = pym.VoxelDomain(nx, ny, nz)
        ndof, bc, f = problem_setup(problem, nx, ny, nz)
        self.f, self.ndof_total = f, f.size
        self.sx = pym.Signal("x", state=np.full(dom.nel, 0.5))
        with pym.Network() as fn:
            sxf = pym.DensityFilter(dom, radius=RADIUS)(self.sx)
            ss = pym.MathExpression(f"{XMIN} + {1.0 - XMIN}*inp0^3")(sxf)
            sK = (pym.AssembleStiffness if ndof == 3 else pym.AssemblePoisson)(dom, bc=bc)(ss)
            mgs = [pym.solvers.GeometricMultigrid(dom)]  # examples/topology_optimization/ex_compliance_multigrid.py:107-121
            while True:
                sub = mgs[-1].sub_domain
                if any(n % 2 != 0 for n in sub.size) or any(sub.size < 8):
                    break
                mgs.append(pym.solvers.GeometricMultigrid(sub))
                mgs[-2].inner_level = mgs[-1]
            self.cg = pym.solvers.CG(preconditioner=mgs[0], tol=TOL)
            self.su = pym.LinSolve(hermitian=True, solver=self.cg)(sK, f)
            self.sc = pym.EinSum("i,i->")(self.su, f)
        self.fn, self.nlevels, self.nel = fn, len(mgs), dom.nel

    def step(self, x, first=False):
        if not first:
            self.sx.state = x
            self.fn.response()
        self.fn.reset()
        self.sc.sensitivity = 1.0
        self.fn.sensitivity()
        return float(self.sc.state), self.sx.sensitivity


def cpu_arm(size, steps, warmup, problem="cantilever"):
    """Time the reference on the host cores.  Falls back to the oracle port only if the reference cannot be imported."""
    pym, root = load_reference()
    t0 = time.perf_counter()
    if pym is not None:
        chain = ReferenceChain(pym, size, problem)  # building the Network evaluates the first design (x = 0.5) once
        kind, where = "reference", root
    else:
        from oracle import Grid
        from oracle.chain import ComplianceProblem

        class _Port:
            def __init__(self):
                kinds = {"cantilever": "cantilever", "mbb": "mbb3d", "thermal": "heatsink"}
                self.P = ComplianceProblem(Grid(*size), kind=kinds[problem], radius=RADIUS, xmin=XMIN, tol=TOL)
                self.nel, self.ndof_total, self.nlevels = self.P.grid.nel, self.P.f.size, len(self.P.mgs)

            def step(self, x, first=False):
                c = self.P.response(x)
                return c, self.P.sensitivity()

        chain, kind, where = _Port(), "port", "oracle/ (reference not importable)"
    setup = time.perf_counter() - t0
    xs = design_sequence(chain.nel, warmup + steps)
    times, compl = [], []
    cpu0 = wall0 = None
    for i, x in enumerate(xs):
        if i == warmup:
            cpu0, wall0 = time.process_time(), time.perf_counter()
        t0 = time.perf_counter()
        c, dx = chain.step(x, first=(i == 0 and kind == "reference"))
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            compl.append(c)
    cores = max(1, round((time.process_time() - cpu0) / max(time.perf_counter() - wall0, 1e-9)))
    return dict(sec_per_iter=sum(times) / len(times), setup_s=setup, ndof=chain.ndof_total, compliance=compl, cores=cores,
                kind=kind, where=where, nlevels=chain.nlevels, times=times)


def pick_cpu_size(full, n_iters, explicit, budget_s=240.0, problem="cantilever"):
    """Largest sample grid whose whole run (set-up + n_iters iterations of the unmodified reference) fits ``budget_s`` on THIS
    host: the table below was measured in the build container (128x64x64: ~11 s / iteration, 25 GB RSS, ~40 s set-up --
    BASELINE.md section 2); a two-iteration calibration run at 32x16x16 (~0.13 s / iteration there) scales it to the host the
    bench runs on.  Returns (size, host speed factor)."""
    if explicit is not None:
        return tuple(explicit), None
    try:
        import psutil

        ram = psutil.virtual_memory().available
    except Exception:
        ram = 0
    factor = 1.0
    try:
        factor = max(0.5, cpu_arm((32, 16, 16), 2, 1, problem)["sec_per_iter"] / 0.13)
    except Exception:
        pass
    cands = [((128, 64, 64), 11.0, 40.0, 34e9), ((96, 48, 48), 4.6, 18.0, 16e9), ((64, 32, 32), 1.3, 6.0, 6e9)]
    for size, per_iter, setup, need in cands:
        if all(s <= f for s, f in zip(size, full)) and ram >= need and factor * (setup + n_iters * per_iter) <= budget_s:
            return size, factor
    return ((64, 32, 32) if min(full) >= 32 else tuple(full)), factor


def cpu_line(r, size, full, problem, steps, warmup):
    ndof_full = (3 if problem != "thermal" else 1) * (full[0] + 1) * (full[1] + 1) * (full[2] + 1)
    scale = r["ndof"] / ndof_full
    value = scale / r["sec_per_iter"]
    sample = (f"{'UNMODIFIED pyMOTO (' + str(r['where']) + ')' if r['kind'] == 'reference' else 'oracle port'}: the same design iteration "
              f"(DensityFilter, MathExpression SIMP, AssembleStiffness, LinSolve(CG(tol 1e-8, GeometricMultigrid x{r['nlevels']})), EinSum, "
              f"backward pass) on a {size[0]}x{size[1]}x{size[2]} grid ({r['ndof']} dof = {scale:.5f} of the full workload), {steps} timed "
              f"iterations after {warmup} warm-up ({r['sec_per_iter']:.3f} s / iteration, one-off setup {r['setup_s']:.1f} s excluded), "
              f"time scaled linearly in dof by 1/{scale:.5f}; {os.cpu_count()} host cores present, {r['cores']} busy on average "
              f"(scipy's csr_matvec / csr_matmat and np.add.at are single-threaded); PARDISO row skipped: mkl not installed")
    return value, scale, sample


def run_reference(args, full):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    size, host_factor = pick_cpu_size(full, args.steps + args.warmup, args.cpu_size, problem=args.problem)
    r = cpu_arm(size, args.steps, args.warmup, args.problem)
    value, scale, sample = cpu_line(r, size, full, args.problem, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * r["sec_per_iter"] / scale, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"3D {args.problem} compliance {full[0]}x{full[1]}x{full[2]} hex8, CG(1e-8)+GMG, DensityFilter r=2",
                   "same_config": False,
                   "caveat": f"the reference needs ~15 kB/dof of host RAM for its set-up (~200 GB at 256x128x128): it is MEASURED on the "
                             f"{size[0]}x{size[1]}x{size[2]} sample and value / ms_per_step are that measurement scaled linearly in dof",
                   "sample_grid": list(size), "measured_sec_per_iter_on_sample": r["sec_per_iter"], "extrapolation_factor": 1.0 / scale,
                   "measured_run_s": r["setup_s"] + sum(r["times"]), "host_speed_vs_build_container": host_factor},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "compliance_on_sample": r["compliance"],
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock / throttle-reason samples DURING the timed region.  Primary: NVML polled in-process every 5 ms (the timed
    region is a few hundred ms, too short for a freshly spawned ``nvidia-smi -lms`` to report anything); fallback: an
    ``nvidia-smi`` poller that must be started before the warm-up.  Only samples stamped inside [begin(), end()] count."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc, self.nvml, self.h = gpu_index, [], None, None, None
        self.t0 = self.t1 = None
        self._stop = threading.Event()
        self.source = None

    def _nvml_handle(self):
        import pynvml

        pynvml.nvmlInit()
        try:
            import torch

            uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in vis.split(",") if v.strip().isdigit()]
            phys = int(ids[self.gpu]) if self.gpu < len(ids) else self.gpu
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)

    def start(self):
        """Call BEFORE the warm-up steps."""
        try:
            self.nvml, self.h = self._nvml_handle()
            self.max_sm = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.h, self.nvml.NVML_CLOCK_SM))
            self.source = "nvml"
            self.t = threading.Thread(target=self._poll_nvml, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.t = threading.Thread(target=self._read_smi, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        n = self.nvml
        bits = ((n.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"), (n.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (n.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"), (n.nvmlClocksEventReasonSwPowerCap, "sw_power_cap"))
        while not self._stop.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                try:
                    mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((time.perf_counter(), sm, self.max_sm, [nm for b, nm in bits if mask & b]))
            except Exception:
                pass
            time.sleep(0.005)

    def _read_smi(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            if len(r) >= 9 and r[1].replace(".", "").isdigit() and r[2].replace(".", "").isdigit():
                self.rows.append((time.perf_counter(), float(r[1]), float(r[2]),
                                  [nm for nm, v in zip(self.NAMES, r[5:9]) if v.lower().startswith("active")]))

    def begin(self):
        self.t0 = time.perf_counter()

    def end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        self._stop.set()
        if self.proc is not None:
            self.proc.terminate()
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi on this host"], "samples": 0}
        self.t.join(timeout=2)
        t0 = self.t0 if self.t0 is not None else -1e300
        t1 = self.t1 if self.t1 is not None else 1e300
        rows = [r for r in self.rows if t0 <= r[0] <= t1]
        scope = "timed region"
        if not rows:  # poller too slow for the region: fall back to everything since start() (warm-up included) and say so
            rows, scope = list(self.rows), "warm-up + timed region"
        reasons = sorted({nm for r in rows for nm in r[3]})
        return {"sm_mhz": statistics.median([r[1] for r in rows]) if rows else None,
                "sm_max_mhz": max(r[2] for r in rows) if rows else None, "reasons": reasons, "samples": len(rows),
                "source": self.source, "scope": scope}


# ------------------------------------------------------------------------------------------------ GPU arm
class GpuChain:
    """The design iteration through the public Module API of pymoto_b200 (device-resident tensors).  With more than one
    rank the grid is split into z-slabs (pymoto_b200/slab.py) and every rank holds its own element layers / node planes."""

    def __init__(self, size, world=1, problem="cantilever", slab_kw=None, min_size=8):
        import pymoto_b200 as pmb
        from pymoto_b200 import device as dv

        self.pmb, self.dv = pmb, dv
        nx, ny, nz = size
        self.dom = dom = pmb.VoxelDomain(nx, ny, nz)
        ndof = 1 if problem == "thermal" else 3
        self.mgs = pmb.solvers.auto_multigrid(dom, min_size=min_size)
        self.ctx = (pmb.slab.init(dom, n_levels=len(self.mgs) + 1, ndof=ndof, **(slab_kw or {})) if world > 1
                    else pmb.slab.context(nz))
        self.k0, self.k1 = self.ctx.part.planes(0) if world > 1 else (0, nz + 1)
        self.e0, self.e1 = self.ctx.part.elem_layers(0) if world > 1 else (0, nz)
        self.lay = nx * ny
        _, bc, f = problem_setup(problem, nx, ny, nz, self.k0, self.k1)
        self.bc, self.ndof = bc, ndof
        self.ndof_global = dom.nnodes * ndof
        self.f = dv.to_device(f)
        self.flt = pmb.DensityFilter(dom, radius=RADIUS)
        self.simp = pmb.SIMP(XMIN, 3)
        self.asm = (pmb.AssemblePoisson if problem == "thermal" else pmb.AssembleStiffness)(dom, bc=bc)
        self.cg = pmb.solvers.CG(preconditioner=self.mgs[0], tol=TOL)
        self.ls = pmb.LinSolve(hermitian=True, solver=self.cg)
        self.compl = pmb.Compliance()

    def local(self, x_global):
        return x_global[self.e0 * self.lay:self.e1 * self.lay]

    def step(self, x):
        y = self.flt(x)
        s = self.simp(y)
        K = self.asm(s)
        u = self.ls(K, self.f)
        c = self.compl(u, self.f)
        du, _ = self.compl._sensitivity(1.0)
        dK, _ = self.ls._sensitivity(du)
        ds = self.asm._sensitivity(dK)[0]
        dy = self.simp._sensitivity(ds)
        dx = self.flt._sensitivity(dy)
        return c, dx


def parity_gate(world, rank):
    """Small-grid gate against the CPU oracle BEFORE anything is timed: one design iteration (seeded random design) of the
    cantilever at 16 x 8 x (8 N) -- slab-decomposed with split multigrid levels when N > 1 -- compared with the numpy / scipy
    restatement of the reference on the whole grid.  Raises if a north-star tolerance is missed."""
    import torch
    import torch.distributed as dist
    import pymoto_b200 as pmb
    from oracle import Grid
    from oracle.chain import ComplianceProblem

    nx, ny, nz = 16, 8, 8 * world
    P = ComplianceProblem(Grid(nx, ny, nz), kind="cantilever", tol=TOL, min_size=4)
    x = np.random.default_rng(5).random(P.grid.nel)
    c_ref = P.response(x)
    dx_ref = P.sensitivity()
    chain = GpuChain((nx, ny, nz), world, "cantilever", slab_kw=dict(min_planes=2, min_dofs=0), min_size=4)
    c, dx = chain.step(chain.dv.to_device(chain.local(x).copy()))
    relres = pmb.solvers.LinearSolver.residual(chain.asm._mat, chain.ls._u_dev, chain.f)
    if world > 1:
        parts = [torch.empty_like(dx) for _ in range(world)]
        dist.all_gather(parts, dx)
        dx = torch.cat(parts)
    dx = dx.cpu().numpy()
    out = {"grid": [nx, ny, nz], "compliance_rel": abs(float(c) - c_ref) / abs(c_ref),
           "dcdx_rel": float(np.abs(dx - dx_ref).max() / np.abs(dx_ref).max()), "relres": float(relres),
           "cg_its": int(chain.cg.iterations), "oracle_its": int(P.cg.iterations), "compliance": float(c), "oracle_compliance": c_ref,
           "split_levels": chain.ctx.part.n_dist if world > 1 else None,
           "tolerances": {"compliance_rel": 1e-6, "dcdx_rel": 1e-6, "relres": 1e-8, "cg_its": "+-1"}}
    out["pass"] = bool(out["compliance_rel"] <= 1e-6 and out["dcdx_rel"] <= 1e-6 and out["relres"] <= 1e-8
                       and abs(out["cg_its"] - out["oracle_its"]) <= 1)
    pmb.slab.reset()
    if not out["pass"]:
        raise SystemExit(f"[bench] parity gate FAILED on rank {rank}: {out}")
    return out


def algorithmic_bytes(stats, mats):
    """SURVEY.md 8d byte model applied to the calls of a timed region: per operator application 8 nnz (stencil-CSR values,
    no index traffic) + vector passes; assembly 8 nnz + 8 nel; Galerkin 8 nnz_fine + 8 nnz_coarse; transfers and vector
    kernels by their operand count.  ``mats``: nx -> (n, nnz, nel)."""
    total = 0.0
    for (name, det), cnt in stats.items():
        if det is None:
            continue
        key, mode = det
        if name in ("pmb_spmv", "pmb_elem_spmv") and key in mats:
            n, nnz, nel = mats[key]
            per = (8 * nnz if name == "pmb_spmv" else 8 * nel) + {0: 16 * n, 1: 24 * n, 2: 32 * n}[mode]
        elif name == "pmb_assemble" and key in mats:
            per = 8 * mats[key][1] + 8 * mats[key][2]
        elif name in ("pmb_galerkin_cols", "pmb_galerkin_rows", "pmb_galerkin_direct") and key in mats:
            per = 4 * mats[key][1] + 0.5 * mats[key][1] / 8 * 8  # half of (read fine + write coarse = nnz/8) each
        elif name == "pmb_rowstats" and key in mats:
            per = 8 * mats[key][1] + 12 * mats[key][0]
        elif name in ("pmb_restrict", "pmb_prolong_add") and key in mats:
            per = (16 if name == "pmb_prolong_add" else 8) * mats[key][0] + mats[key][0]
        elif name in ("pmb_lincomb", "pmb_cg_xr_update"):
            per = 24 * key
        elif name in ("pmb_dots", "pmb_smooth0", "pmb_mask_zero", "pmb_bc_split", "pmb_vec_div", "pmb_simp", "pmb_simp_bwd"):
            per = 16 * key
        elif name == "pmb_assemble_sens" and key in mats:
            per = 16 * mats[key][0] + 8 * mats[key][2]
        elif name == "pmb_filter_apply" and key in mats:
            per = 16 * mats[key][2]
        else:
            continue
        total += per * cnt
    return total


def run_b200(args, full):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # keep stdout for the ONE JSON line: anything libraries print (e.g. the NCCL version banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as ge

    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from pymoto_b200 import _lib, device as dv
    from pymoto_b200.matrix import DeviceCSR
    import pymoto_b200 as pmb

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxreduce(v):
        if world == 1:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.csr:
        DeviceCSR.matrix_free = False
    parity = None if (args.no_parity or args.kernel_only) else parity_gate(world, rank)

    chain = GpuChain(full, world, args.problem)
    W, K = args.warmup, args.steps
    xs_host = design_sequence(chain.dom.nel, W + K + (1 if args.profile else 0), keep=chain.local)
    nel = xs_host[0].size  # elements held by this rank
    pinned = [torch.from_numpy(x).pin_memory() for x in xs_host]
    xs_dev = [p.to("cuda", non_blocking=True) for p in pinned]
    torch.cuda.synchronize()

    if args.kernel_only:  # assemble once, then operator applications on the finest level only
        K0 = chain.asm(chain.simp(chain.flt(xs_dev[0])))
        D0 = K0.diagonal_device()
        va, vb, vc = K0.new_vec(zero=True), K0.new_vec(zero=True), dv.to_device(np.random.default_rng(0).random(K0.shape[0]))
        for _ in range(12):
            K0.apply(_lib.JACOBI, va, vb, b=vc, diag=D0, w=0.5)
            va, vb = vb, va
        torch.cuda.synchronize()
        return
    # ---------------- device-resident timing
    sampler = ClockSampler(local)
    sampler.start()
    for i in range(W):
        chain.step(xs_dev[i])
    barrier()
    # the host drives ~1000 launches and one poll per CG iteration per step: a generation-2 collection of the interpreter in the
    # middle of a step (tens of ms with torch / numpy / scipy loaded) would be charged to the step -- collect now, then keep
    # the collector off inside the timed regions (what timeit does); nothing is allocated cyclically by the path itself
    import gc
    gc.collect()
    gc.disable()
    sampler.begin()
    launches0, stats0 = _lib.launch_count, dict(_lib.call_stats)
    comm0 = (chain.ctx.comm.exchanges, chain.ctx.comm.allreduces, getattr(chain.ctx.comm, "fast_allreduces", 0))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    cg_its, compl = [], []
    ev[0].record()
    for i in range(K):
        c, dx = chain.step(xs_dev[W + i])
        ev[i + 1].record()
        cg_its.append(chain.cg.iterations)
        compl.append(c)
    barrier()
    gc.enable()
    sampler.end()
    clocks = sampler.stop()
    if world > 1:
        chain.ctx.comm.check_peer_timeouts()  # a one-launch exchange that gave up waiting would make the timing meaningless
    launches = _lib.launch_count - launches0
    stats = {k: v - stats0.get(k, 0) for k, v in _lib.call_stats.items() if v - stats0.get(k, 0) > 0}
    comm_counts = {"halo_exchanges_per_step": (chain.ctx.comm.exchanges - comm0[0]) / K,
                   "allreduces_per_step": (chain.ctx.comm.allreduces - comm0[1]) / K,
                   "launches_per_step": launches / K,
                   "one_launch_peer_exchange": bool(getattr(chain.ctx.comm, "fused", False)),
                   "one_launch_peer_allreduces_per_step": (chain.ctx.comm.fast_allreduces - comm0[2]) / K} if world > 1 else None
    total_ms = maxreduce(ev[0].elapsed_time(ev[K]))
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(K)]
    # one slab-decomposed job over all ranks.  Weak scaling: the grid grows with the rank count, so the whole-job
    # throughput is quoted in 256x128x128-equivalent design iterations (iterations/s x dof / 12.83 M dof)
    iters_per_sec = K / (total_ms * 1e-3)
    dof_scale = chain.ndof_global / (3 * 257 * 129 * 129) if (world > 1 and args.problem != "thermal") else 1.0
    value = iters_per_sec * dof_scale
    compl = [float(c) for c in compl]
    step_avg_ms = total_ms / K

    if args.profile:
        from pymoto_b200.solvers import GeometricMultigrid as _GMG

        graphs, _GMG.use_cuda_graph = _GMG.use_cuda_graph, False  # time every call individually
        _lib.profile_times = {}
        chain.step(xs_dev[W + K])
        prof, _lib.profile_times = _lib.profile_times, None
        _GMG.use_cuda_graph = graphs
        tot = sum(v[1] for v in prof.values())
        if rank == 0:
            print(f"# per-call breakdown of one step (synchronised calls), total {tot:.2f} ms", file=sys.stderr)
            for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1]):
                print(f"# {v[1]:9.3f} ms {100 * v[1] / tot:5.1f}%  n={v[0]:5d}  avg {1e3 * v[1] / v[0]:9.1f} us  {k}", file=sys.stderr)

    # ---------------- end to end from pinned host buffers (x in, compliance + dc/dx out)
    e2e = None
    if not args.no_e2e:
        chain.ls._u_dev = None  # same cold-start state as the device-resident run's first warm-up step
        out_pinned = torch.empty(nel, dtype=torch.float64).pin_memory()
        c_pinned = torch.empty(1, dtype=torch.float64).pin_memory()
        xd = torch.empty(nel, dtype=torch.float64, device="cuda")  # the user's device copy of the design, filled from the host every step
        for i in range(W):
            xd.copy_(pinned[i], non_blocking=True)
            chain.step(xd)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2e_steps = []
        gc.collect()
        gc.disable()
        dev_allocs0 = torch.cuda.memory_stats().get("num_device_alloc", 0)
        e0.record()
        for i in range(K):
            t_step = time.perf_counter()
            xd.copy_(pinned[W + i], non_blocking=True)
            c, dx = chain.step(xd)
            out_pinned.copy_(dx, non_blocking=True)
            c_pinned.copy_(c.reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()  # the user reads the result on the host every step
            e2e_steps.append(1e3 * (time.perf_counter() - t_step))
        e1.record()
        gc.enable()
        barrier()
        ms = maxreduce(e0.elapsed_time(e1))
        e2e = {"value": dof_scale * K / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * nel * world,
               "d2h_bytes_per_step": (8 * nel + 8) * world, "ms_per_step_list": e2e_steps,
               # cudaMalloc calls of the caching allocator inside the timed region (each one synchronises the device: a step that
               # contains one shows up as an outlier in ms_per_step_list)
               "device_allocs_in_timed_region": torch.cuda.memory_stats().get("num_device_alloc", 0) - dev_allocs0}

    # ---------------- roofline of the kernels that carry the step
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        hbm_peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    import ctypes as C

    probe_out = dv.empty(_lib.query("pmb_probe_fp64_out_doubles"))
    fp64 = {}
    for kind, nm in ((0, "dfma"), (1, "dmma")):
        tf = C.c_double(0.0)
        _lib.call("pmb_probe_fp64", kind, 4000, dv.ptr(probe_out), C.byref(tf), dv.stream())
        fp64[nm] = float(tf.value)
    fp64_peak = max(fp64.values())

    A = chain.asm._mat
    n, nnz, ndof_ = A.shape[0], A.nnz, A.grid.ndof
    mg0 = chain.mgs[0]
    D = mg0.smoother.D

    def time_sweeps(op, u, u2, b, Dg, reps=20):
        for _ in range(3):
            op.apply(_lib.JACOBI, u, u2, b=b, diag=Dg, w=0.5)
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(reps):
            op.apply(_lib.JACOBI, u, u2, b=b, diag=Dg, w=0.5)
            u, u2 = u2, u
        k1.record()
        torch.cuda.synchronize()
        return k0.elapsed_time(k1) / reps

    def sum_calls(name, nx=None):
        return sum(v for (nm, det), v in stats.items() if nm == name and det is not None and (nx is None or det[0] == nx))

    was_mf = DeviceCSR.matrix_free and A.generator is not None
    fine_mf, fine_csr = sum_calls("pmb_elem_spmv", full[0]), sum_calls("pmb_spmv", full[0])
    bufs0 = (mg0._buf["u"], mg0._buf["u2"], mg0._buf["t"])
    # stencil-CSR kernel on the finest ASSEMBLED matrix (in the step only with --csr / in the csr_streamed leg)
    DeviceCSR.matrix_free = False
    csr0_ms = time_sweeps(A, *bufs0, D)
    DeviceCSR.matrix_free = was_mf or DeviceCSR.matrix_free
    if not args.csr:
        DeviceCSR.matrix_free = True
    csr0_bytes = 8 * nnz + 32 * n
    csr_level0 = {"kernel": f"tile_kernel<{ndof_},JACOBI> on the finest assembled matrix", "kernel_ms": csr0_ms,
                  "algorithmic_bytes": csr0_bytes, "achieved": csr0_bytes / (csr0_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                  "frac": csr0_bytes / (csr0_ms * 1e-3) / 1e9 / hbm_peak, "launches_per_step": fine_csr / K,
                  "in_timed_step": bool(args.csr), "reference_layout_gbs": (12 * nnz + 4 * (n + 1) + 40 * n) / (csr0_ms * 1e-3) / 1e9}
    csr_level1 = None
    if len(chain.mgs) > 1 and chain.mgs[0].Ac is not None and chain.mgs[1]._buf is not None:
        A1, mg1 = chain.mgs[0].Ac, chain.mgs[1]
        l1_ms = time_sweeps(A1, mg1._buf["u"], mg1._buf["u2"], mg1._buf["t"], mg1.smoother.D)
        l1_bytes = 8 * A1.nnz + 32 * A1.shape[0]
        l1_calls = sum_calls("pmb_spmv", full[0] // 2)
        csr_level1 = {"kernel": f"tile_kernel<{ndof_},JACOBI> on the level-1 Galerkin operator", "kernel_ms": l1_ms,
                      "algorithmic_bytes": l1_bytes, "achieved": l1_bytes / (l1_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                      "frac": l1_bytes / (l1_ms * 1e-3) / 1e9 / hbm_peak, "launches_per_step": l1_calls / K,
                      "share_of_step": (l1_calls / K) * l1_ms / step_avg_ms}
    traffic, traffic_src = None, None
    try:  # DRAM bytes of the dominant kernel from the committed `ncu --set full` capture of the same build, if there is one
        prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_dominant_kernel_r2.json")))
        if tuple(prof.get("grid", [])) == tuple(full) and world == 1 and prof.get("problem") == args.problem and not args.csr:
            traffic, traffic_src = prof["dram_bytes_read"] + prof["dram_bytes_write"], prof.get("source")
    except Exception:
        pass
    if was_mf and not args.csr:
        mf_ms = time_sweeps(A, *bufs0, D)
        gen = A.generator
        dense_flops = 2.0 * 8 * (8 * ndof_ * ndof_ + ndof_) * (n / ndof_)  # 8 elements x (8 nodes x ndof^2 + ndof) FMA per node
        if gen.variant >= 8 and ndof_ in (1, 3):
            # parity-block layouts (pmb_elem_par.cuh): per element 8 blocks of ndof x ndof (2 flop per entry) + density scaling
            # (8 ndof) + butterflies: forward 16 ndof (x / y of one plane 8 ndof, z 8 ndof), transposed 20 ndof (z 8 ndof, carry
            # 4 ndof, y / x 8 ndof) + 3 ndof to gather the node + ~4 ndof epilogue; one element per node
            flops = (2.0 * 8 * ndof_ * ndof_ + (8 + 16 + 20 + 3 + 4) * ndof_) * (n / ndof_)
        else:
            flops = dense_flops
        mf_bytes = 40 * n + 8 * nel
        tf = flops / (mf_ms * 1e-3) / 1e12
        hbm_view = {"algorithmic_bytes": mf_bytes, "achieved": mf_bytes / (mf_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": mf_bytes / (mf_ms * 1e-3) / 1e9 / hbm_peak}
        common = {"kernel": f"matrix-free finest-level operator, layout {gen.variant} of pmb_elem_spmv (Jacobi sweep from the element densities)",
                  "kernel_ms": mf_ms, "launches_per_step": fine_mf / K, "share_of_step": (fine_mf / K) * mf_ms / step_avg_ms,
                  "flops_per_launch": flops, "dense_product_flops_per_launch": dense_flops,
                  "dense_equivalent_tflops": dense_flops / (mf_ms * 1e-3) / 1e12, "fp64_probe_tflops": fp64, "layout_ms_autotune": DeviceCSR.elem_timings_ms.get(ndof_),
                  "traffic": traffic, "traffic_source": traffic_src, "peak_source": "pmb_probe_fp64 in this run (register-only DFMA / DMMA streams)",
                  "hbm_peak_source": peak_src}
        # arithmetic intensity flops / byte vs the ridge fp64_peak / hbm_peak decides which roof bounds the kernel
        if flops / mf_bytes >= fp64_peak * 1e12 / (hbm_peak * 1e9):
            roofline = {"bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak, "hbm": hbm_view, **common}
        else:
            roofline = {"bound": "hbm", **hbm_view, "fp64": {"achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak}, **common}
        roofline["speedup_vs_streaming_assembled_values"] = csr0_ms / mf_ms
    else:
        roofline = {"bound": "hbm", **{k: csr_level0[k] for k in ("achieved", "peak", "unit", "frac", "kernel", "kernel_ms", "algorithmic_bytes")},
                    "traffic": None, "launches_per_step": fine_csr / K, "share_of_step": (fine_csr / K) * csr0_ms / step_avg_ms,
                    "peak_source": peak_src}
    roofline["csr_level1"], roofline["csr_level0"] = csr_level1, csr_level0

    # ---------------- the same K steps with every level streamed from its assembled CSR values (north-star layout) and the
    #                  whole iteration against SURVEY 8d's byte model
    csr_streamed = None
    if was_mf and not args.csr and world == 1 and not args.no_e2e:
        DeviceCSR.matrix_free = False
        chain.ls._u_dev = None
        for i in range(W):
            chain.step(xs_dev[i])
        torch.cuda.synchronize()
        s0 = dict(_lib.call_stats)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for i in range(K):
            chain.step(xs_dev[W + i])
        c1.record()
        torch.cuda.synchronize()
        DeviceCSR.matrix_free = True
        cstats = {k: v - s0.get(k, 0) for k, v in _lib.call_stats.items() if v - s0.get(k, 0) > 0}
        cms = c0.elapsed_time(c1) / K
        mats, lvl = {}, chain.mgs[0]
        Al = A
        while Al is not None:
            g_ = Al.grid
            mats[g_.nx] = (Al.shape[0], Al.nnz, g_.nx * g_.ny * max(g_.nz, 1))
            Al, lvl = (lvl.Ac, lvl.inner_level) if isinstance(lvl, pmb.solvers.GeometricMultigrid) else (None, None)
        it_bytes = algorithmic_bytes(cstats, mats) / K
        csr_streamed = {"value": 1e3 / cms, "unit": UNIT, "ms_per_step": cms,
                        "note": "same workload with DeviceCSR.matrix_free = False (bench.py --csr): every level streams 8 B / non-zero"}
        roofline["iteration"] = {"leg": "csr_streamed", "algorithmic_bytes_per_step": it_bytes, "ms_per_step": cms,
                                 "achieved": it_bytes / (cms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": it_bytes / (cms * 1e-3) / 1e9 / hbm_peak,
                                 "model": "SURVEY.md 8d formulas on the calls counted in the timed region (8 B / nnz layout)"}

    # ---------------- BASELINE configs[3] / configs[4] beside the metric config (short runs, same process)
    secondary = {}
    if not args.no_secondary and args.problem == "cantilever" and args.size is None:
        def side_run(size, problem, note):
            pmb.slab.reset()
            ch = GpuChain(size, world, problem)
            xs = [dv.to_device(x) for x in design_sequence(ch.dom.nel, 3 + 5, keep=ch.local)]
            for i in range(3):
                ch.step(xs[i])
            barrier()
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            its, cs = [], []
            q0.record()
            for i in range(5):
                c_, _ = ch.step(xs[3 + i])
                its.append(ch.cg.iterations)
                cs.append(c_)
            q1.record()
            barrier()
            ms = maxreduce(q0.elapsed_time(q1)) / 5
            return {"workload": note, "grid": list(size), "dof": ch.ndof_global, "n_gpus": world, "ms_per_step": ms,
                    "iters_per_sec": 1e3 / ms, "cg_iterations": its, "compliance": [float(c_) for c_ in cs], "steps": 5, "warmup": 3}

        del xs_dev
        torch.cuda.empty_cache()
        if world > 1:
            secondary["configs3_mbb"] = side_run(full, "mbb", "3D half-MBB compliance (SURVEY 8d bc set), z-slab partitioned")
        if world in (1, 8):
            secondary["configs4_thermal"] = side_run((256, 256, 256), "thermal", "3D heat-sink scalar conduction 256^3, CG+GMG")

    # ---------------- CPU baseline (bounded sample of the same workload, unmodified reference), rank 0 only
    cpu = None
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        size = tuple(args.cpu_size) if args.cpu_size else ((64, 32, 32) if min(full) >= 32 else tuple(full))
        r = cpu_arm(size, 3, 1, args.problem)
        v, scale, sample = cpu_line(r, size, full, args.problem, 3, 1)
        cpu = {"value": v, "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": sample, "extrapolation_factor": 1.0 / scale}

    if rank == 0:
        gen = A.generator
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": step_avg_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"3D {args.problem} compliance {full[0]}x{full[1]}x{full[2]} hex8 ({chain.ndof_global} dof, "
                                   f"{n} dof / nnz {nnz} per GPU), "
                                   f"SIMP p=3 xmin=1e-9, DensityFilter r=2, LDAS+CG(tol 1e-8)+GMG({len(chain.mgs)} levels, "
                                   "5+5 Jacobi w=0.5, V-cycle replayed as one CUDA graph on 1 GPU), warm start, seeded design perturbations; "
                                   "finest-level operator " +
                                   (f"matrix-free (layout {gen.variant})" if was_mf and not args.csr else "streamed from the assembled CSR values"),
                       "l2": "inputs larger than L2 (8.2 GB of matrix values written per step, > 100 MB vectors per operator application)",
                       "parallelism": "1 GPU" if world == 1 else
                       f"{world} z-slabs ({chain.ctx.part.n_dist} split multigrid levels, coarser levels replicated), " +
                       ("halo exchange and dot-product all-reduce as one-launch kernels over NVLink peer memory (pmb_peer_*)"
                        if getattr(chain.ctx.comm, "fused", False) else "peer-memory halo mailboxes + NCCL all-reduce") +
                       "; value = iterations/s x dof / 12.83M (weak scaling: a NORMALISED number, "
                       "iters_per_sec_this_grid is the raw rate)",
                       "iters_per_sec_this_grid": iters_per_sec},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "parity": parity, "comm": comm_counts,
            "csr_streamed": csr_streamed, "cpu_baseline": cpu, "secondary": secondary or None,
            "cg_iterations": cg_its, "ms_per_step_list": step_ms, "compliance": compl,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.size:
        full = tuple(args.size)
    elif args.problem == "thermal":
        full = (256, 256, 256)
    else:
        full = WEAK_GRIDS.get(world, (256, 128, 128 * world))
    if args.impl == "reference":
        run_reference(args, full)
    else:
        run_b200(args, full)


if __name__ == "__main__":
    main()
